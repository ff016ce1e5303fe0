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
//                 tcgen05.ld -> BN statistics (warp transpose-reduce, then shared-memory accumulators that persist
//                 across the CTA's tiles and are flushed with one atomic per channel when the column block changes)
//                 -> bf16 -> XOR-swizzled staging tile -> 128-byte-contiguous global stores.  The TMEM buffer is
//                 released right after the tcgen05.ld, before the global stores.
#pragma once

namespace seg {
namespace tc {

constexpr int V2_BN = 128;
constexpr int V2_STAGES = 5;
constexpr int V2_STAGE_BYTES = A_BYTES + V2_BN * 128;  // 32 KB
constexpr int V2_THREADS = 320;
constexpr int V2_STAGING_BYTES = BM * V2_BN * 2;       // 32 KB
constexpr int V2_SMEM = V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + 1024 /*stat accumulators*/ + 256 /*barriers*/ + 1024 /*align*/;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(V2_THREADS, 1) conv_gemm_tc2(const __grid_constant__ TcParams p) {
  constexpr int BN = V2_BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem0 = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - raw_addr);
  uint8_t* stage = smem_gen + V2_STAGES * V2_STAGE_BYTES;                       // staging tile [128][256 B]
  float* stat_sm = reinterpret_cast<float*>(stage + V2_STAGING_BYTES);          // [2][128]
  const uint32_t bar0 = smem0 + V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + 1024;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (V2_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * V2_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * V2_STAGES + 2 + b); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * (2 * V2_STAGES + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + V2_STAGES * V2_STAGE_BYTES + V2_STAGING_BYTES + 1024 + 8 * (2 * V2_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.Ncols + BN - 1) / BN;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int num_tiles = m_tiles * n_tiles;
  const int iters_per_tile = p.taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
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
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) stat_sm[threadIdx.x - 64] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

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
    }
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    const int e = warp - 2;
    const int lg = warp & 3;  // TMEM lane group
    const int h = e >> 2;     // column half of the tile
    const int trow = lg * 32 + lane;
    const int etid = e * 32 + lane;  // 0..255
    int cur_n = -1;
    int i = 0;
    float v[32];
    auto flush_stats = [&](int n_tile) {
      // called by all 256 epilogue threads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (n_tile >= 0) {
        const int c = etid & 127, which = etid >> 7;
        const int col = n_tile * BN + c;
        const float val = stat_sm[which * BN + c];
        if (col < p.Ncols && val != 0.f) atomicAdd(p.stats + (size_t)which * p.Ncols + col, val);
        stat_sm[which * BN + c] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      const int n_tile = t % n_tiles;
      const int m0 = (t / n_tiles) * BM, n0 = n_tile * BN;
      const int row = m0 + trow;
      const bool row_ok = row < p.M;
      if (p.stats && n_tile != cur_n) {
        flush_stats(cur_n);
        cur_n = n_tile;
      }
      mbar_wait(tfull_bar(b), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      float st1[2], st2[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ch = h * 2 + c;  // 32-column chunk of the tile
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + ch * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c16 = ch * 4 + g;
          *reinterpret_cast<bf16x8*>(stage + (size_t)trow * (BN * 2) + ((c16 ^ (trow & 15)) << 4)) = pack8(v + g * 8);
        }
        st1[c] = st2[c] = 0.f;
        if (p.stats) {
          float s1[32], s2[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float x = row_ok ? v[q] : 0.f;
            s1[q] = x;
            s2[q] = x * x;
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float send1 = up ? s1[j] : s1[j + off];
              const float keep1 = up ? s1[j + off] : s1[j];
              s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
              const float send2 = up ? s2[j] : s2[j + off];
              const float keep2 = up ? s2[j + off] : s2[j];
              s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
            }
          }
          st1[c] = s1[0];
          st2[c] = s2[0];
        }
      }
      // accumulator fully read: hand the TMEM buffer back to the MMA warp before touching global memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));
      if (p.stats) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = (h * 2 + c) * 32 + lane;
          atomicAdd(stat_sm + col, st1[c]);
          atomicAdd(stat_sm + BN + col, st2[c]);
        }
      }
      // stream this warp's 32 rows x 64 columns out: 8 lanes cover one 128-byte row segment, 4 rows per instruction
      const int ncols_tile = min(BN, p.Ncols - n0);
      const bool vec_ok = ((p.ldo * 2) & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.out) + (size_t)n0 * 2) & 15) == 0;
#pragma unroll 2
      for (int idx = lane; idx < 32 * 8; idx += 32) {
        const int rr = idx >> 3;
        const int c16 = h * 8 + (idx & 7);
        const int first_col = c16 * 8;
        if (first_col >= ncols_tile) continue;
        const int tr = lg * 32 + rr;
        const int grow = m0 + tr;
        if (grow >= p.M) continue;
        long long pixel = grow;
        if (p.out_strided) {
          const int n = grow / p.PQ;
          const int rem = grow - n * p.PQ;
          const int ii = rem / p.Q, jj = rem - ii * p.Q;
          pixel = ((long long)n * p.out_H + (ii * p.osy + p.opy)) * p.out_W + (jj * p.osx + p.opx);
        }
        bf16x8 val = *reinterpret_cast<const bf16x8*>(stage + (size_t)tr * (BN * 2) + ((c16 ^ (tr & 15)) << 4));
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)pixel * p.ldo + n0 + first_col;
        if (vec_ok && first_col + 8 <= ncols_tile) {
          if (p.beta != 0.f) {
            float a[8], bb[8];
            unpack8(val, a);
            unpack8(*reinterpret_cast<const bf16x8*>(dst), bb);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += p.beta * bb[k];
            val = pack8(a);
          }
          *reinterpret_cast<bf16x8*>(dst) = val;
        } else {
          float a[8];
          unpack8(val, a);
          for (int k = 0; k < 8; ++k)
            if (first_col + k < ncols_tile) {
              float o = a[k];
              if (p.beta != 0.f) o += p.beta * bf2f(dst[k]);
              dst[k] = f2bf(o);
            }
        }
      }
      __syncwarp();  // the staging region of this warp is reused by its next tile
    }
    if (p.stats) flush_stats(cur_n);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int KIND>
static int launch_v2(const TcParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  auto kfn = conv_gemm_tc2<KIND>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM);
    SEG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(v2 smem=%d): %s", V2_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int n_tiles = ceil_div(p.Ncols, V2_BN);
  const int64_t num_tiles = ceil_div64(p.M, BM) * n_tiles;
  int grid = num_sms();
  // a grid that is a multiple of the number of column blocks pins every CTA to one column block: the weight tile stays
  // hot and the BN-statistics accumulators are flushed once per CTA instead of once per tile
  if (n_tiles <= grid) grid = (grid / n_tiles) * n_tiles;
  if ((int64_t)grid > num_tiles) grid = (int)num_tiles;
  kfn<<<grid, V2_THREADS, V2_SMEM, stream>>>(p);
  return check_launch("conv_gemm_tc2");
}

}  // namespace tc
}  // namespace seg
