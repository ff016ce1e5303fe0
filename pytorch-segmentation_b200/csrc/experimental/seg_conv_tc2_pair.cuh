// seg_conv_tc2_pair.cuh — DRAFT of the CTA-pair (cta_group::2) variant of the persistent tcgen05 implicit-GEMM convolution
// (DESIGN.md §9.1).  Written without GPU access at the end of round 1: it COMPILES for sm_100a (tools/check_pair_ptx.sh builds
// it through experimental/pair_kernel_check.cu) and has NEVER RUN.  It is not included by the library.  To bring it up:
// include it at the bottom of seg_conv_tc.cu, route eligible shapes to launch_pair<KIND> behind an environment switch, and
// compare against conv_gemm_tc2 with tools/conv_micro.py (every mbarrier wait is bounded, so a protocol bug traps instead of
// hanging the box).
//
// Why: conv_gemm_tc2 moves 32 KB of operands from L2 into each SM per 128x128x64 MMA step — 128 B/flop-unit the L2 cannot
// feed at full tensor rate for the 1x1 layers (25-45 % tensor-pipe).  Here a cluster of two CTAs owns a 256 x 256 tile: each
// CTA loads its own 128 A rows (16 KB) and HALF of the 256 B rows/columns (16 KB) per k-step and the leader issues ONE
// tcgen05.mma.cta_group::2 with M = 256, N = 256 that reads both halves from both SMs: the same 32 KB per CTA now feed twice
// the flops.  TMEM: 2 accumulator buffers x 256 columns = all 512 columns of each SM.
//
//   cluster (2,1,1), rank 0 = leader; 320 threads per CTA as in conv_gemm_tc2:
//   warp 0      : TMA producer in BOTH CTAs (A rows m0 + 128*rank, B half 128*rank); transaction bytes of both land on the
//                 leader's full barrier (cta_group::2 loads, peer-masked barrier address)
//   warp 1      : TMEM pair allocation in both CTAs; MMA issue + commits (multicast to both CTAs) in the leader only
//   warps 2..9  : epilogue in both CTAs: each CTA drains ITS 128 accumulator rows x 256 columns (two 64-column slices per
//                 warp) through per-warp swizzled staging boxes and TMA stores / bf16 reduce-adds; BN statistics as in v2
// Barriers (protocol in seg_ptx_pair.cuh): full[s] count 1 (leader's is the one that counts), empty[s] count 1 per CTA
// (released by commit-multicast), tfull[b] count 1 per CTA (commit-multicast), tempty[b] count 16 on the leader (8 epilogue
// warps of each CTA arrive remotely).
#pragma once
#include "seg_ptx_pair.cuh"

namespace seg {
namespace tc {

using namespace ptx2;

constexpr int P2_BN = 256;                                   // columns of the pair tile (UMMA N)
constexpr int P2_HALF = 128;                                 // B rows / columns each CTA loads
constexpr int P2_STAGES = 4;
constexpr int P2_STAGE_BYTES = A_BYTES + P2_HALF * 128;      // 32 KB per CTA and k-step
constexpr int P2_THREADS = 320;
constexpr int P2_STAGING_BYTES = BM * P2_BN * 2;             // 64 KB: 16 boxes of [32 rows][64 columns] bf16
constexpr int P2_STAT_BYTES = 4 * 2 * P2_BN * 4;             // [lane group][sum, sum of squares][column] fp32
constexpr int P2_SMEM = P2_STAGES * P2_STAGE_BYTES + P2_STAGING_BYTES + P2_STAT_BYTES + 256 /*barriers*/ + 1024 /*align*/;
static_assert(P2_SMEM <= 227 * 1024, "pair kernel shared memory exceeds the per-CTA limit");

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1) conv_gemm_tc2_pair(const __grid_constant__ TcParams p) {
  constexpr int BN = P2_BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem0 = (raw_addr + 1023u) & ~1023u;  // identical in both CTAs (same kernel, same dynamic smem base)
  uint8_t* smem_gen = smem_raw + (smem0 - raw_addr);
  uint8_t* stage = smem_gen + P2_STAGES * P2_STAGE_BYTES;                       // 16 staging boxes of 4 KB
  float* stat_sm = reinterpret_cast<float*>(stage + P2_STAGING_BYTES);          // [4][2][256]
  const uint32_t bar0 = smem0 + P2_STAGES * P2_STAGE_BYTES + P2_STAGING_BYTES + P2_STAT_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (P2_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * P2_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * P2_STAGES + 2 + b); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * (2 * P2_STAGES + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + P2_STAGES * P2_STAGE_BYTES + P2_STAGING_BYTES + P2_STAT_BYTES + 8 * (2 * P2_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_tiles = (p.Ncols + BN - 1) / BN;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);  // 256-row pair tiles
  const int num_tiles = m_tiles * n_tiles;
  const int iters_per_tile = p.taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    prefetch_tmap(&p.mapC);
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is waited on)
    }
    fence_mbar_init();
  }
  if (warp == 1) {  // the same warp index in both CTAs, the same result slot offset
    tmem2_alloc(tmem_ptr_addr, 2 * BN);
    tmem2_relinquish();
  }
  for (int i = threadIdx.x; i < P2_STAT_BYTES / 4; i += P2_THREADS) stat_sm[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers are initialised before anything signals them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (lane == 0) {
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters) {
        const int m0 = (t / n_tiles) * (2 * BM) + rank * BM;
        const int n0 = (t % n_tiles) * BN + rank * P2_HALF;  // this CTA's half of the B rows (KK) / columns (KM)
        int n_img = 0, h0 = 0, w0 = 0;
        if (p.x_im2col) row_to_coords(m0, p.PQ, p.Q, p.stride, p.lower_h, p.lower_w, n_img, h0, w0);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int wtap = p.tap_wt[tap];
          const uint16_t oh = (uint16_t)p.tap_oh[tap], ow = (uint16_t)p.tap_ow[tap];
          for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
            const int st = it % P2_STAGES;
            const uint32_t ph = (it / P2_STAGES) & 1;
            mbar_wait(empty_bar(st), ph ^ 1u);  // my own slot is free (the leader's commit releases both CTAs)
            const uint32_t a_dst = smem0 + st * P2_STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(full_bar(st), 2 * P2_STAGE_BYTES);  // bytes of BOTH CTAs
            if (p.x_im2col)
              tma2_load_im2col_4d(a_dst, &p.mapA, full_bar(st), kc * BK, w0, h0, n_img, ow, oh);
            else
              tma2_load_2d(a_dst, &p.mapA, full_bar(st), kc * BK, m0);
            if (KIND == KIND_KK) {
              tma2_load_2d(b_dst, &p.mapB, full_bar(st), kc * BK, wtap * p.brows_per_tap + n0);  // box [128 rows][64 k]
            } else {
#pragma unroll
              for (int j = 0; j < P2_HALF / 64; ++j)  // boxes [64 k-rows][64 cols]
                tma2_load_2d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, wtap * p.brows_per_tap + kc * BK);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (lane == 0 && rank == 0) {
      constexpr int B_MN = (KIND == KIND_KK) ? 0 : 1;
      constexpr uint32_t idesc = make_idesc_bf16_m256(BN, 0, B_MN);
      int it = 0, i = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters, ++i) {
        const int b = i & 1;
        mbar_wait(tempty_bar(b), (((uint32_t)i >> 1) & 1u) ^ 1u);  // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(b * BN);
        for (int j = 0; j < iters_per_tile; ++j, ++it) {
          const int st = it % P2_STAGES;
          const uint32_t ph = (it / P2_STAGES) & 1;
          mbar_wait(full_bar(st), ph);  // both CTAs' tiles of this stage have landed
          tc_fence_after();
          const uint32_t a_addr = smem0 + st * P2_STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_BYTES;
          const uint64_t adesc0 = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = adesc0 + (uint64_t)(k * 32 >> 4);
            const uint64_t bdesc = bdesc0 + (uint64_t)(B_MN ? (k * 2048 >> 4) : (k * 32 >> 4));
            umma2_bf16(tacc, adesc, bdesc, idesc, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma2_commit_multicast(empty_bar(st), (uint16_t)0x3);  // frees the slot in BOTH CTAs
        }
        umma2_commit_multicast(tfull_bar(b), (uint16_t)0x3);     // accumulator halves are complete in BOTH CTAs
      }
    }
  } else {
    // =============================== epilogue (warps 2..9, both CTAs) ===============================
    const int e = warp - 2;
    const int lg = warp & 3;  // TMEM lane group (a warp may only touch lanes 32*(warp%4)..)
    const int hh = e >> 2;    // 0 / 1: this warp drains column slices hh and hh + 2 (64 columns each)
    const int etid = e * 32 + lane;  // 0..255
    int cur_n = -1;
    int i = 0;
    float v[32], v1[32];
    auto flush_stats = [&](int n_tile) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (n_tile >= 0) {
        const int c = etid;  // one column per epilogue thread, both moments
        const int col = n_tile * BN + c;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          float val = 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            val += stat_sm[(g * 2 + which) * BN + c];
            stat_sm[(g * 2 + which) * BN + c] = 0.f;
          }
          if (col < p.Ncols && val != 0.f) atomicAdd(p.stats + (size_t)which * p.Ncols + col, val);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int t = cluster_id; t < num_tiles; t += n_clusters, ++i) {
      const int b = i & 1;
      const int n_tile = t % n_tiles;
      const int m0 = (t / n_tiles) * (2 * BM) + rank * BM, n0 = n_tile * BN;
      if (p.stats && n_tile != cur_n) {
        flush_stats(cur_n);
        cur_n = n_tile;
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int slice = hh + 2 * s2;  // 64-column slice of the 256-column tile
        uint8_t* region = stage + slice * 16384 + lg * 4096;  // [32 rows][128 B], SWIZZLE_128B pattern, this warp only
        if (lane == 0) bulk_wait_group_read0();  // my previous boxes have been read out of shared memory
        __syncwarp();
        if (s2 == 0) {
          mbar_wait(tfull_bar(b), ((uint32_t)i >> 1) & 1u);
          tc_fence_after();
        }
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + slice * 64), v);
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + slice * 64 + 32), v1);
        tmem_ld_wait();
        if (s2 == 1) {  // both slices of this warp are in registers: hand the buffer back (leader's barrier, from either CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta0(tempty_bar(b));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<bf16x8*>(region + lane * 128 + ((g ^ (lane & 7)) << 4)) = pack8(v + g * 8);
          *reinterpret_cast<bf16x8*>(region + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) = pack8(v1 + g * 8);
        }
        __syncwarp();
        if (p.stats)
          warp_column_stats<128, 7>(region + (lane & 3) * 4, (uint32_t)(lane >> 2) << 4, min(32, p.M - m0 - lg * 32),
                                    stat_sm + (size_t)(lg * 2) * BN + slice * 64 + (lane >> 2) * 8 + (lane & 3) * 2, BN);
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0 && m0 + lg * 32 < p.M && n0 + slice * 64 < p.Ncols) {
          if (p.beta != 0.f)
            tma_reduce_add_2d(&p.mapC, smem_u32(region), n0 + slice * 64, m0 + lg * 32);  // dst += tile, bf16 add in L2
          else
            tma_store_2d(&p.mapC, smem_u32(region), n0 + slice * 64, m0 + lg * 32);
          bulk_commit_group();
        }
      }
    }
    if (p.stats) flush_stats(cur_n);
    if (lane == 0) bulk_wait_group0();  // all boxes written before the CTA retires
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync();  // neither CTA leaves (or frees TMEM) while its peer can still signal into it
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, 2 * BN);
  }
}

// Host side.  Requirements beyond conv_gemm_tc2's TMA-epilogue path: the caller built mapB with 128-row boxes for KK (as for
// V2_BN = 128) / 64-row boxes for KM, beta is 0 or 1, no strided output.  Grid = 2 x clusters, at most one CTA per SM.
template <int KIND>
static int launch_pair(TcParams p, cudaStream_t stream) {
  static bool attr_set = false;
  auto kfn = conv_gemm_tc2_pair<KIND>;
  SEG_REQUIRE(!p.out_strided && p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (p.beta == 0.f || p.beta == 1.f),
              "conv_gemm_tc2_pair: needs the TMA epilogue (dense, 16-byte aligned output, beta 0 or 1)");
  SEG_REQUIRE(!(p.beta != 0.f && p.stats != nullptr), "conv_gemm_tc2_pair: beta-accumulate and BN statistics are not combined");
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
    SEG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(pair smem=%d): %s", P2_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  if (make_map_2d(&p.mapC, p.out, p.M, p.Ncols, p.ldo, 32)) return 1;
  const int n_tiles = ceil_div(p.Ncols, P2_BN);
  const int64_t num_tiles = ceil_div64(p.M, 2 * BM) * n_tiles;
  int clusters = num_sms() / 2;
  if (n_tiles <= clusters) clusters = (clusters / n_tiles) * n_tiles;  // every cluster stays on one column block
  if ((int64_t)clusters > num_tiles) clusters = (int)num_tiles;
  kfn<<<dim3(2 * clusters), dim3(P2_THREADS), (size_t)P2_SMEM, stream>>>(p);
  return check_launch("conv_gemm_tc2_pair");
}

}  // namespace tc
}  // namespace seg
