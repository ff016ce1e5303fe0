// seg_conv_tc2.cuh — persistent variant of the tcgen05 implicit-GEMM convolution (fprop KK / dgrad KM, bf16 output).
// Included by seg_conv_tc.cu (shares TcParams, tensor-map builders and the PTX wrappers).
//
// One CTA per SM loops over 128x128 output tiles (static round-robin schedule).  Two TMEM accumulator buffers
// (2 x 128 columns) let the single-thread MMA issuer start tile i+1 while the EIGHT epilogue warps drain tile i;
// the 5-stage TMA ring keeps filling across tile boundaries.  This is what the short-K pointwise convolutions need
// (1x1 convs with C = 64..256 have only 1-4 k-iterations per tile: a one-tile-per-CTA kernel is all prologue and
// epilogue).  Per-CTA setup (barrier init, TMEM allocation, descriptor prefetch) is paid once per launch.
//   warp 0      : TMA producer (im2col-mode / 2-D activation tiles + weight tiles)
//   warp 1      : TMEM alloc (256 cols) + tcgen05.mma issue + tcgen05.commit
//   warps 2..9  : epilogue; warp (lg = warp%4, h = (warp-2)/4) owns TMEM lanes 32*lg.. and columns 64*h.. of the tile:
//                 tcgen05.ld -> bf16 -> XOR-swizzled staging tile -> [BN statistics: column sums of the staged rows
//                 into shared-memory accumulators that persist across the CTA's tiles and are flushed with one atomic
//                 per channel when the column block changes] -> 128-byte-contiguous global stores.  The TMEM buffer
//                 is released right after the tcgen05.ld, before the global stores.  BETA variant (dgrad accumulating
//                 into an existing gradient): the destination is prefetched before the accumulator wait.
#pragma once

namespace seg {
namespace tc {

constexpr int V2_BN = 128;     // default tile width
constexpr int V2_THREADS = 320;
// Tile-width dependent layout.  BNT = 128: 5 stages of 32 KB.  BNT = 256 (TMA epilogue only): 3 stages of 48 KB, both TMEM
// accumulators = all 512 columns; a 128 x 256 tile moves 48 KB of operands per 64-deep k-step for twice the MACs of a
// 128 x 128 tile's 32 KB — 0.75x the L2->SM bytes per flop, which is what bounds these kernels.
template <int BNT>
struct V2Cfg {
  static constexpr int STAGES = (BNT == 128) ? 5 : 3;
  static constexpr int STAGE_BYTES = A_BYTES + BNT * 128;
  static constexpr int STAGING_BYTES = BM * 128 * 2;      // 32 KB: eight [32 rows][64 columns] boxes (one pass of 128 columns)
  static constexpr int STAT_BYTES = 4 * 2 * BNT * 4;      // [lane group][sum, sum of squares][column] fp32
  static constexpr int SMEM = STAGES * STAGE_BYTES + STAGING_BYTES + STAT_BYTES + 256 /*barriers*/ + 1024 /*align*/;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int EPI_MANUAL = 0;       // per-thread 16-byte global stores (strided sub-grid outputs, unaligned outputs)
constexpr int EPI_MANUAL_BETA = 1;  // the same, accumulating into the destination (prefetched)
constexpr int EPI_TMA = 2;          // per-warp TMA store / bf16 reduce-add of a [32 rows][64 columns] box

// Column sums / sums of squares of 32 staged bf16 rows x 64 columns by one warp: lane = 2 adjacent columns.
// `base` = first row of the warp's region (+ 4 * (lane & 3)), `cx` = un-swizzled byte offset of the lane's 16-byte chunk.
template <int ROW_BYTES, int SWZ_MASK>
__device__ __forceinline__ void warp_column_stats(const uint8_t* base, uint32_t cx, int rmax, float* slot, int slot_stride) {
  float a0[2] = {0.f, 0.f}, a1[2] = {0.f, 0.f}, q0[2] = {0.f, 0.f}, q1[2] = {0.f, 0.f};
  if (rmax >= 32) {
#pragma unroll
    for (int r = 0; r < 32; ++r) {  // row offsets and swizzle masks are immediates
      const uint32_t w = *reinterpret_cast<const uint32_t*>(base + r * ROW_BYTES + (cx ^ ((uint32_t)(r & SWZ_MASK) << 4)));
      const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
      a0[r & 1] += x0;
      a1[r & 1] += x1;
      q0[r & 1] = fmaf(x0, x0, q0[r & 1]);
      q1[r & 1] = fmaf(x1, x1, q1[r & 1]);
    }
  } else {
    for (int r = 0; r < rmax; ++r) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(base + r * ROW_BYTES + (cx ^ ((uint32_t)(r & SWZ_MASK) << 4)));
      const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
      a0[0] += x0;
      a1[0] += x1;
      q0[0] = fmaf(x0, x0, q0[0]);
      q1[0] = fmaf(x1, x1, q1[0]);
    }
  }
  // the calling warp is the only writer of its (lane group, column) slots: plain read-modify-write, no atomics
  float2 s1 = *reinterpret_cast<float2*>(slot), s2 = *reinterpret_cast<float2*>(slot + slot_stride);
  s1.x += a0[0] + a0[1];
  s1.y += a1[0] + a1[1];
  s2.x += q0[0] + q0[1];
  s2.y += q1[0] + q1[1];
  *reinterpret_cast<float2*>(slot) = s1;
  *reinterpret_cast<float2*>(slot + slot_stride) = s2;
}

template <int KIND, int EPI, int BNT = V2_BN>
__global__ void __launch_bounds__(V2_THREADS, 1) conv_gemm_tc2(const __grid_constant__ TcParams p) {
  constexpr int BN = BNT;
  static_assert(BNT == 128 || (BNT == 256 && EPI == EPI_TMA), "256-wide tiles use the TMA epilogue");
  using VC = V2Cfg<BNT>;
  constexpr int V2_STAGES = VC::STAGES, V2_STAGE_BYTES = VC::STAGE_BYTES, V2_STAGING_BYTES = VC::STAGING_BYTES, V2_STAT_BYTES = VC::STAT_BYTES;
  constexpr int NPASS = BNT / 128;  // epilogue passes of 128 columns over a tile
  constexpr bool BETA = (EPI == EPI_MANUAL_BETA);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem0 = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - raw_addr);
  uint8_t* stage = smem_gen + V2_STAGES * V2_STAGE_BYTES;                       // staging tile [128][256 B]
  float* stat_sm = reinterpret_cast<float*>(stage + V2_STAGING_BYTES);          // [4][2][128]
  const uint32_t bar0 = smem0 + V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + V2_STAT_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (V2_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * V2_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * V2_STAGES + 2 + b); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * (2 * V2_STAGES + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + V2_STAT_BYTES + 8 * (2 * V2_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.Ncols + BN - 1) / BN;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int num_tiles = m_tiles * n_tiles;
  const int iters_per_tile = p.taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    if (EPI == EPI_TMA) prefetch_tmap(&p.mapC);
    for (int s = 0; s < V2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, 2 * BN);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < V2_STAT_BYTES / 4; i += V2_THREADS) stat_sm[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();  // setup above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        int n_img = 0, h0 = 0, w0 = 0;
        if (p.x_im2col) row_to_coords(m0, p.PQ, p.Q, p.stride, p.lower_h, p.lower_w, n_img, h0, w0);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int wtap = p.tap_wt[tap];
          const uint16_t oh = (uint16_t)p.tap_oh[tap], ow = (uint16_t)p.tap_ow[tap];
          for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
            const int st = it % V2_STAGES;
            const uint32_t ph = (it / V2_STAGES) & 1;
            mbar_wait(empty_bar(st), ph ^ 1u);
            const uint32_t a_dst = smem0 + st * V2_STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            mbar_arrive_expect_tx(full_bar(st), V2_STAGE_BYTES);
            if (p.x_im2col)
              tma_load_im2col_4d(a_dst, &p.mapA, full_bar(st), kc * BK, w0, h0, n_img, ow, oh);
            else
              tma_load_2d(a_dst, &p.mapA, full_bar(st), kc * BK, m0);
            if (KIND == KIND_KK) {
              tma_load_2d(b_dst, &p.mapB, full_bar(st), kc * BK, wtap * p.brows_per_tap + n0);  // box [128][64]
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)  // boxes [64 k-rows][64 cols]
                tma_load_2d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, wtap * p.brows_per_tap + kc * BK);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr int B_MN = (KIND == KIND_KK) ? 0 : 1;
      constexpr uint32_t idesc = make_idesc_bf16(BN, 0, B_MN);
      int it = 0, i = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
        const int b = i & 1;
        mbar_wait(tempty_bar(b), (((uint32_t)i >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(b * BN);
        for (int j = 0; j < iters_per_tile; ++j, ++it) {
          const int st = it % V2_STAGES;
          const uint32_t ph = (it / V2_STAGES) & 1;
          mbar_wait(full_bar(st), ph);
          tc_fence_after();
          const uint32_t a_addr = smem0 + st * V2_STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_BYTES;
          const uint64_t adesc0 = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = adesc0 + (uint64_t)(k * 32 >> 4);
            const uint64_t bdesc = bdesc0 + (uint64_t)(B_MN ? (k * 2048 >> 4) : (k * 32 >> 4));
            umma_bf16(tacc, adesc, bdesc, idesc, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(st));
        }
        umma_commit(tfull_bar(b));
      }
      pdl_trigger();  // this CTA's loads and MMAs are all issued: the next kernel's CTAs may take their places now
    }
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    const int e = warp - 2;
    const int lg = warp & 3;  // TMEM lane group
    const int h = e >> 2;     // column half of the tile
    const int trow = lg * 32 + lane;
    const int etid = e * 32 + lane;  // 0..255
    int i = 0;
    float v[32];
    // BN statistics: the host sizes the grid as a multiple of the number of column blocks, so this CTA stays on column
    // block blockIdx.x % n_tiles for all of its tiles and its shared-memory accumulators are flushed ONCE, at the end, with
    // one fp64 atomic per channel (exact accumulation of fp32 partials: see conv_gemm_tc's epilogue — bit-reproducible)
    auto finish_stats = [&]() {
      // called by all 256 epilogue threads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int my_n = blockIdx.x % n_tiles;
      for (int idx = etid; idx < 2 * BN; idx += 256) {
        const int which = idx / BN, c = idx - which * BN;
        const int col = my_n * BN + c;
        float val = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) val += stat_sm[(g * 2 + which) * BN + c];
        if (col < p.Ncols) atomicAdd(p.stats + (size_t)which * p.Ncols + col, (double)val);
      }
      if (p.sync.world > 0) {  // SyncBN: the last CTA pushes the finished totals to every peer and raises the flags
        volatile int* flag = reinterpret_cast<volatile int*>(smem_gen + V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + V2_STAT_BYTES + 128);
        sync_push_when_last(p.sync, p.stats, 2 * p.Ncols, p.stat_ticket, gridDim.x, etid, 256,
                            [] { asm volatile("bar.sync 1, 256;" ::: "memory"); }, flag);
      }
    };
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      const int n_tile = t % n_tiles;
      const int m0 = (t / n_tiles) * BM, n0 = n_tile * BN;
      if constexpr (EPI == EPI_TMA) {
        // ---- TMA epilogue: per pass of 128 columns this warp owns rows 32*lg.. and columns 64*h.. = one [32][64] box whose
        //      staging region (4 KB, 128-byte rows, SWIZZLE_128B pattern) only this warp touches ----
        uint8_t* region = stage + h * 16384 + lg * 4096;
        mbar_wait(tfull_bar(b), ((uint32_t)i >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          const int cbase = ps * 128 + h * 64;  // first column of this warp's box inside the tile
          if (lane == 0) bulk_wait_group_read0();  // my previous box has been read out of shared memory
          __syncwarp();
          float v1[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + cbase), v);
          tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + cbase + 32), v1);
          tmem_ld_wait();
          if (ps == NPASS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(b));  // accumulator drained: the MMA warp may reuse it
          }
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              *reinterpret_cast<bf16x8*>(region + lane * 128 + ((g ^ (lane & 7)) << 4)) = pack8(v + g * 8);
              *reinterpret_cast<bf16x8*>(region + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) = pack8(v1 + g * 8);
            }
          }
          __syncwarp();
          if (p.stats && !(p.dbg & 2))
            warp_column_stats<128, 7>(region + (lane & 3) * 4, (uint32_t)(lane >> 2) << 4, min(32, p.M - m0 - lg * 32),
                                      stat_sm + (size_t)(lg * 2) * BN + cbase + (lane >> 2) * 8 + (lane & 3) * 2, BN);
          fence_proxy_async();  // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && !(p.dbg & 1) && m0 + lg * 32 < p.M && n0 + cbase < p.Ncols) {
            if (p.beta != 0.f)
              tma_reduce_add_2d(&p.mapC, smem_u32(region), n0 + cbase, m0 + lg * 32);  // dst += tile, bf16 add in L2
            else
              tma_store_2d(&p.mapC, smem_u32(region), n0 + cbase, m0 + lg * 32);
            bulk_commit_group();
          }
        }
        continue;
      }
      // beta-accumulate (dgrad into a gradient that already holds the skip branch): fetch this lane's eight 16-byte
      // pieces of the destination NOW — eight loads in flight per thread while the accumulator is still being
      // produced, instead of one dependent load per store (which left the epilogue bound by DRAM latency)
      const int ncols_tile = min(BN, p.Ncols - n0);
      const bool vec_ok = ((p.ldo * 2) & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.out) + (size_t)n0 * 2) & 15) == 0;
      const int c16 = h * 8 + (lane & 7);
      const int first_col = c16 * 8;
      const bool col_ok = first_col < ncols_tile;
      const bool vec = vec_ok && first_col + 8 <= ncols_tile;
      long long off[8];
      bf16x8 resid[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int grow = m0 + lg * 32 + (lane >> 3) + 4 * k;
        off[k] = -1;
        if (col_ok && grow < p.M) {
          long long pixel = grow;
          if (p.out_strided) {
            const int n = grow / p.PQ;
            const int rem = grow - n * p.PQ;
            const int ii = rem / p.Q, jj = rem - ii * p.Q;
            pixel = ((long long)n * p.out_H + (ii * p.osy + p.opy)) * p.out_W + (jj * p.osx + p.opx);
          }
          off[k] = pixel * p.ldo + n0 + first_col;
          if (BETA && vec && !(p.dbg & 1)) resid[k] = *reinterpret_cast<const bf16x8*>(reinterpret_cast<const __nv_bfloat16*>(p.out) + off[k]);
        }
      }
      mbar_wait(tfull_bar(b), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (p.dbg & 4) break;
        const int ch = h * 2 + c;  // 32-column chunk of the tile
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + ch * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cc = ch * 4 + g;
          *reinterpret_cast<bf16x8*>(stage + (size_t)trow * (BN * 2) + ((cc ^ (trow & 15)) << 4)) = pack8(v + g * 8);
        }
      }
      // accumulator fully read: hand the TMEM buffer back to the MMA warp before touching global memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));
      if (!BETA && p.stats && !(p.dbg & 2)) {
        // BN statistics of the tile as STORED (bf16-rounded, i.e. exactly what bn_apply will normalise): every lane
        // sums two adjacent columns over this warp's 32 staged rows — conflict-free 4-byte shared loads, no shuffles
        warp_column_stats<BN * 2, 15>(stage + (size_t)(lg * 32) * (BN * 2) + (lane & 3) * 4, (uint32_t)(h * 8 + (lane >> 2)) << 4,
                                      min(32, p.M - m0 - lg * 32),
                                      stat_sm + (size_t)(lg * 2) * BN + h * 64 + (lane >> 2) * 8 + (lane & 3) * 2, BN);
      }
      // stream this warp's 32 rows x 64 columns out: 8 lanes cover one 128-byte row segment, 4 rows per instruction
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (off[k] < 0 || (p.dbg & 1)) continue;
        const int tr = lg * 32 + (lane >> 3) + 4 * k;
        bf16x8 val = *reinterpret_cast<const bf16x8*>(stage + (size_t)tr * (BN * 2) + ((c16 ^ (tr & 15)) << 4));
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off[k];
        if (vec) {
          if (BETA) {
            float a[8], bb[8];
            unpack8(val, a);
            unpack8(resid[k], bb);
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] += p.beta * bb[q];
            val = pack8(a);
          }
          *reinterpret_cast<bf16x8*>(dst) = val;
        } else {
          float a[8];
          unpack8(val, a);
          for (int q = 0; q < 8; ++q)
            if (first_col + q < ncols_tile) {
              float o = a[q];
              if (BETA) o += p.beta * bf2f(dst[q]);
              dst[q] = f2bf(o);
            }
        }
      }
      __syncwarp();  // the staging region of this warp is reused by its next tile
    }
    if (EPI != EPI_MANUAL_BETA && p.stats) finish_stats();
    if (EPI == EPI_TMA && lane == 0) bulk_wait_group0();  // all boxes written before the CTA retires
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int KIND, int EPI, int BNT = V2_BN>
static int launch_v2_impl(const TcParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  auto kfn = conv_gemm_tc2<KIND, EPI, BNT>;
  constexpr int SMEM = V2Cfg<BNT>::SMEM;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SEG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(v2 smem=%d): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int n_tiles = ceil_div(p.Ncols, BNT);
  const int64_t num_tiles = ceil_div64(p.M, BM) * n_tiles;
  int grid = num_sms();
  // a grid that is a multiple of the number of column blocks pins every CTA to one column block: the weight tile stays
  // hot and the BN-statistics accumulators are flushed once per CTA instead of once per tile
  if (n_tiles <= grid) grid = (grid / n_tiles) * n_tiles;
  if ((int64_t)grid > num_tiles) grid = (int)num_tiles;
  SEG_REQUIRE(p.stats == nullptr || grid % n_tiles == 0, "conv_gemm_tc2: statistics need a grid that pins every CTA to one column block");
  launch_pdl(kfn, dim3(grid), dim3(V2_THREADS), (size_t)SMEM, stream, p);
  return check_launch("conv_gemm_tc2");
}

// `p` by value: the destination map and the experiment flags are filled in here
template <int KIND>
static int launch_v2(TcParams p, cudaStream_t stream, int bnt = V2_BN) {
  p.dbg = env_dbg();
  const bool tma_ok = !p.out_strided && p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                      (p.beta == 0.f || p.beta == 1.f) && !(p.dbg & 16);
  if (tma_ok) {
    if (make_map_2d(&p.mapC, p.out, p.M, p.Ncols, p.ldo, 32)) return 1;
    if (bnt == 256) return launch_v2_impl<KIND, EPI_TMA, 256>(p, stream);
    return launch_v2_impl<KIND, EPI_TMA>(p, stream);
  }
  SEG_REQUIRE(bnt == V2_BN, "conv_gemm_tc2: 256-wide tiles need the TMA epilogue");
  if (p.beta != 0.f) {
    SEG_REQUIRE(p.stats == nullptr, "conv_gemm_tc2: beta-accumulate and BN statistics are not combined on this path");
    return launch_v2_impl<KIND, EPI_MANUAL_BETA>(p, stream);
  }
  return launch_v2_impl<KIND, EPI_MANUAL>(p, stream);
}

// can the persistent kernel take this output through its TMA epilogue (the only epilogue of the 256-wide tiles)?
static bool v2_tma_epilogue_ok(const void* out, long long ldo, float beta, bool strided) {
  return !strided && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (beta == 0.f || beta == 1.f) && !(env_dbg() & 16);
}

}  // namespace tc
}  // namespace seg
