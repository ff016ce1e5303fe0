// seg_conv_tc.cu — tcgen05 + TMA implicit-GEMM convolution for sm_100a (forward, data-gradient, weight-gradient).
//
// Replaces what the reference dispatches to cuDNN for every nn.Conv2d on the hot path (dense 3x3 / dilated 3x3 /
// 1x1): models/deeplabv3_plus.py:256 (ASPP d=6/12/18), :312-315 (decoder 3x3), torchvision
// Bottleneck conv1/conv2/conv3 (layer4 conv2 dilated at deeplabv3_plus.py:47-53), models/resnet.py:43-48.
//
// One CTA computes a 128 x BN fp32 tile held in TMEM.  Warp roles (192 threads):
//   warp 0  : TMA producer.  The activation operand is fetched with an IM2COL-mode tensor map: one instruction
//             brings 128 output pixels x 64 channels of one filter tap (tap displacement passed as the im2col
//             offset, padding = the bounding-box corners, out-of-image taps zero-filled by the TMA unit) straight into
//             the 128B-swizzled K-major layout tcgen05 consumes — the im2col gather happens in the copy engine, never
//             in HBM.  Weights come through a tiled 2-D map over the packed [tap][K][C] matrix.
//   warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per 64-wide k-block).
//   warps 2-5: epilogue; tcgen05.ld 32 lanes x 32 columns -> optional bias -> staged through (now idle) pipeline
//             shared memory with an XOR swizzle -> fully coalesced 16-byte global stores (optionally beta-accumulate,
//             optionally onto a strided sub-grid of the output) ; per-channel sum / sum-of-squares of the fp32
//             accumulators (BatchNorm statistics) reduced per warp by a 31-shuffle transpose-reduce, then per CTA in
//             shared memory, one atomicAdd per channel per CTA.
// Operand-major variants of the same pipeline:
//   KK (fprop) : A = activations (K-major),  B = weights  [tap*K + k][c]   (K-major)
//   KM (dgrad) : A = dY im2col  (K-major),  B = weights  [tap*K + k][c]   (MN-major: c contiguous).  Stride 1: taps
//                flipped.  Stride s > 1: the input pixels are split into s*s parity classes; each class is a stride-1
//                gather over dY with only the taps whose parity matches, written to its strided sub-grid of dX —
//                no wasted MACs on inserted zeros.
//   MM (wgrad) : A = dY [pixel][k] (MN-major), B = X im2col [pixel][c] (MN-major), contraction over pixels,
//                split-K over pixel blocks with fp32 atomics into dW[tap][k][c].
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include "seg_common.cuh"
#include "seg_ptx.cuh"
#include "seg_sync.cuh"

namespace seg {
namespace tc {

using namespace ptx;

constexpr int BM = 128;
constexpr int BK = 64;  // elements per k-block = 128 bytes of bf16
constexpr int A_BYTES = BM * 128;
constexpr int KIND_KK = 0, KIND_KM = 1, KIND_MM = 2;
constexpr int NTHREADS = 192;
constexpr int MAXT = 49;

struct TcParams {
  CUtensorMap mapA;  // KK/KM: activation-side operand ; MM: dY 2-D
  CUtensorMap mapB;  // KK/KM: packed weights 2-D      ; MM: X (im2col or 2-D)
  int M;             // valid output rows
  int Ncols;         // valid output cols
  int taps;          // entries of the tap table
  int kchunks;       // KK/KM: 64-wide channel chunks per tap
  int stride;        // traversal stride of the im2col map (row -> base coordinate)
  int lower_h, lower_w;
  int PQ, Q;         // row index -> (n, p, q)
  int x_im2col;      // activation operand uses the im2col map
  int brows_per_tap; // weight-matrix rows per tap
  short tap_wt[MAXT];  // weight tap index
  short tap_oh[MAXT];  // im2col offset (h)
  short tap_ow[MAXT];  // im2col offset (w)
  void* out;
  long long ldo;
  int out_dtype;
  float beta;
  const float* bias;
  double* stats;     // [2*Ncols] or null, ZERO at launch: per-channel sum / sum of squares of the output (BatchNorm batch
                     // statistics), accumulated with fp64 atomics (see the epilogue)
  unsigned* stat_ticket;   // SyncBN only: one zeroed word counting finished CTAs
  SyncDesc sync;           // world > 0: SyncBN — the last CTA pushes the finished totals to every peer (seg_sync.cuh)
  // strided sub-grid output (stride>1 dgrad): row (n,i,j) -> pixel (n, i*osy+opy, j*osx+opx) of an out_H x out_W map
  int out_strided, out_H, out_W, osy, osx, opy, opx;
  // MM only
  int kblocks_total, kblocks_per_split;
  float* dw;
  int dw_K, dw_C;
  CUtensorMap mapC;  // v2 TMA epilogue: the destination as a 2-D map with [32 rows][64 columns] boxes
  int dbg;  // tuning experiments only (env SEG_TC_DBG): 1 skip global stores/loads, 2 skip statistics, 4 skip TMEM drain
};

template <int BN, int STAGES_>
struct Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void row_to_coords(int m, int PQ, int Q, int stride, int lower_h, int lower_w, int& n, int& h,
                                              int& w) {
  n = m / PQ;
  int rem = m - n * PQ;
  int pp = rem / Q;
  int qq = rem - pp * Q;
  h = lower_h + pp * stride;
  w = lower_w + qq * stride;
}

template <int BN, int STAGES, int KIND>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_tc(const __grid_constant__ TcParams p) {
  using C = Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem0 = (raw_addr + 1023u) & ~1023u;  // 1024-B alignment for SWIZZLE_128B atoms
  uint8_t* smem_gen = smem_raw + (smem0 - raw_addr);
  const uint32_t bar0 = smem0 + C::STAGES * C::STAGE_BYTES;
  // barrier block: full[STAGES], empty[STAGES], tmem_full, then tmem base pointer
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * C::STAGES);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 1));
  const uint32_t tmem_ptr_addr = bar0 + 8u * (2 * C::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- iteration space of the k loop ----
  int tap_mm = 0, kb_begin = 0, kb_end = 0, num_iters;
  if (KIND == KIND_MM) {
    tap_mm = blockIdx.z % p.taps;
    const int split = blockIdx.z / p.taps;
    kb_begin = split * p.kblocks_per_split;
    kb_end = min(kb_begin + p.kblocks_per_split, p.kblocks_total);
    num_iters = max(kb_end - kb_begin, 0);
  } else {
    num_iters = p.taps * p.kchunks;
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();  // setup above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0 && num_iters > 0) {
      if (KIND != KIND_MM) {
        int n_img = 0, h0 = 0, w0 = 0;
        if (p.x_im2col) row_to_coords(m0, p.PQ, p.Q, p.stride, p.lower_h, p.lower_w, n_img, h0, w0);
        int it = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int wtap = p.tap_wt[tap];
          const uint16_t oh = (uint16_t)p.tap_oh[tap], ow = (uint16_t)p.tap_ow[tap];
          for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
            const int st = it % C::STAGES;
            const uint32_t ph = (it / C::STAGES) & 1;
            mbar_wait(empty_bar(st), ph ^ 1u);
            const uint32_t a_dst = smem0 + st * C::STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            mbar_arrive_expect_tx(full_bar(st), C::STAGE_BYTES);
            if (p.x_im2col)
              tma_load_im2col_4d(a_dst, &p.mapA, full_bar(st), kc * BK, w0, h0, n_img, ow, oh);
            else
              tma_load_2d(a_dst, &p.mapA, full_bar(st), kc * BK, m0);
            if (KIND == KIND_KK) {
              tma_load_2d(b_dst, &p.mapB, full_bar(st), kc * BK, wtap * p.brows_per_tap + n0);  // box [BN][64]
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)  // boxes [64 k-rows][64 cols]
                tma_load_2d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, wtap * p.brows_per_tap + kc * BK);
            }
          }
        }
      } else {
        const uint16_t oh = (uint16_t)p.tap_oh[tap_mm], ow = (uint16_t)p.tap_ow[tap_mm];
        int it = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int st = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(empty_bar(st), ph ^ 1u);
          const uint32_t a_dst = smem0 + st * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int pix0 = kb * BK;
          mbar_arrive_expect_tx(full_bar(st), C::STAGE_BYTES);
          tma_load_2d(a_dst, &p.mapA, full_bar(st), m0, pix0);  // dY box [64 pixels][64 k]
          tma_load_2d(a_dst + 8192, &p.mapA, full_bar(st), m0 + 64, pix0);
          if (p.x_im2col) {
            int n_img, h0, w0;
            row_to_coords(pix0, p.PQ, p.Q, p.stride, p.lower_h, p.lower_w, n_img, h0, w0);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_im2col_4d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, w0, h0, n_img, ow, oh);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, pix0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && num_iters > 0) {
      constexpr int A_MN = (KIND == KIND_MM) ? 1 : 0;
      constexpr int B_MN = (KIND == KIND_KK) ? 0 : 1;
      constexpr uint32_t idesc = make_idesc_bf16(BN, A_MN, B_MN);
      for (int it = 0; it < num_iters; ++it) {
        const int st = it % C::STAGES;
        const uint32_t ph = (it / C::STAGES) & 1;
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t a_addr = smem0 + st * C::STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_BYTES;
        // K-major: 8-row atoms 1024 B apart, K advance = 32 B inside the swizzle atom.
        // MN-major: 64-wide MN blocks one box (8192 B) apart, 8-row K groups 1024 B apart, K advance = 16 rows.
        const uint64_t adesc0 = A_MN ? make_smem_desc_sw128(a_addr, 8192, 1024) : make_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adesc = adesc0 + (uint64_t)(A_MN ? (k * 2048 >> 4) : (k * 32 >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(B_MN ? (k * 2048 >> 4) : (k * 32 >> 4));
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(st));  // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
      pdl_trigger();               // loads and MMAs of this CTA are issued: dependents may be scheduled
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int trow = lg * 32 + lane;
    const int row = m0 + trow;
    if (num_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    float v[32];
    if (KIND == KIND_MM) {
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = n0 + ch * 32;
        if (col0 >= p.Ncols) break;  // warp-uniform
        if (num_iters > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
          if (row < p.M) {
            float* dst = p.dw + ((size_t)tap_mm * p.dw_K + row) * p.dw_C + col0;
            if (col0 + 32 <= p.Ncols && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
              // 128-bit vector reductions: 8 RED.F32x4 instead of 32 scalar atomics per thread
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                             "f"(v[i + 2]), "f"(v[i + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.Ncols) atomicAdd(dst + i, v[i]);
            }
          }
        }
      }
    } else {
      // -------- phase 1: TMEM -> registers -> (bias, stats) -> swizzled staging tile in the idle pipeline smem ------
      const bool out_bf16 = (p.out_dtype == SEG_DT_BF16);
      const int esize = out_bf16 ? 2 : 4;
      const int CPR = BN * esize / 16;  // 16-byte chunks per staged row
      const int swz = (CPR >= 32 ? 31 : CPR - 1);
      uint8_t* stage = smem_gen;  // 128 rows x BN*esize bytes
      float* stat_sm = reinterpret_cast<float*>(smem_gen + BM * BN * esize);  // [4 warps][2][BN], right after the staging tile
      const bool row_ok = row < p.M;
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = n0 + ch * 32;
        if (col0 >= p.Ncols) break;  // warp-uniform
        if (num_iters > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.Ncols) v[i] += __ldg(p.bias + col0 + i);
        }
        if (out_bf16) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c16 = ch * 4 + g;
            *reinterpret_cast<bf16x8*>(stage + (size_t)trow * (BN * 2) + ((c16 ^ (trow & swz)) << 4)) = pack8(v + g * 8);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int c16 = ch * 8 + g;
            *reinterpret_cast<float4*>(stage + (size_t)trow * (BN * 4) + ((c16 ^ (trow & swz)) << 4)) =
                make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          }
        }
        if (p.stats) {
          // column sums over this warp's 32 rows: transpose-reduce, lane i ends with column (col0+i)
          float s1[32], s2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // statistics of the value AS STORED (bf16-rounded) — exactly what bn_apply normalises, and what the
            // persistent kernel sums from its staged tile
            const float x = row_ok ? (out_bf16 ? bf2f(f2bf(v[i])) : v[i]) : 0.f;
            s1[i] = x;
            s2[i] = x * x;
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
          stat_sm[(lg * 2 + 0) * BN + ch * 32 + lane] = s1[0];
          stat_sm[(lg * 2 + 1) * BN + ch * 32 + lane] = s2[0];
        }
      }
      __syncwarp();
      // -------- phase 2: each warp streams ITS 32 rows out with coalesced 16-byte stores --------
      const int ncols_tile = min(BN, p.Ncols - n0);           // valid columns of this tile
      const int nchunks = (ncols_tile * esize + 15) >> 4;      // 16-byte chunks that hold valid data
      const bool vec_ok = ((p.ldo * esize) & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.out) + (size_t)n0 * esize) & 15) == 0;
      for (int idx = lane; idx < 32 * CPR; idx += 32) {
        const int rr = idx / CPR;
        const int c16 = idx - rr * CPR;
        if (c16 >= nchunks) continue;
        const int tr = lg * 32 + rr;
        const int grow = m0 + tr;
        if (grow >= p.M) continue;
        long long pixel = grow;
        if (p.out_strided) {
          const int n = grow / p.PQ;
          const int rem = grow - n * p.PQ;
          const int i = rem / p.Q, j = rem - i * p.Q;
          pixel = ((long long)n * p.out_H + (i * p.osy + p.opy)) * p.out_W + (j * p.osx + p.opx);
        }
        const uint8_t* src = stage + (size_t)tr * (BN * esize) + ((c16 ^ (tr & swz)) << 4);
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)pixel * p.ldo + n0) * esize + ((size_t)c16 << 4);
        const int first_col = (c16 << 4) / esize;
        const bool full = first_col + 16 / esize <= ncols_tile;
        if (out_bf16) {
          bf16x8 val = *reinterpret_cast<const bf16x8*>(src);
          if (vec_ok && full) {
            if (p.beta != 0.f) {
              float a[8], b[8];
              unpack8(val, a);
              unpack8(*reinterpret_cast<const bf16x8*>(dst), b);
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] += p.beta * b[k];
              val = pack8(a);
            }
            *reinterpret_cast<bf16x8*>(dst) = val;
          } else {
            float a[8];
            unpack8(val, a);
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
            for (int k = 0; k < 8; ++k)
              if (first_col + k < ncols_tile) {
                float o = a[k];
                if (p.beta != 0.f) o += p.beta * bf2f(d[k]);
                d[k] = f2bf(o);
              }
          }
        } else {
          const float4 val = *reinterpret_cast<const float4*>(src);
          const float a[4] = {val.x, val.y, val.z, val.w};
          float* d = reinterpret_cast<float*>(dst);
          if (vec_ok && full && p.beta == 0.f) {
            *reinterpret_cast<float4*>(d) = val;
          } else {
            for (int k = 0; k < 4; ++k)
              if (first_col + k < ncols_tile) d[k] = (p.beta != 0.f) ? a[k] + p.beta * d[k] : a[k];
          }
        }
      }
      // -------- phase 3: BN statistics: this CTA's column sums (fixed order inside the CTA) are added to the layer's totals
      //          with fp64 atomics.  Each contribution is an fp32 value, so the fp64 sum of the <= few thousand partials of a
      //          channel is EXACT (no rounding at all) whenever their exponents span less than 2^17 — then the order of the
      //          atomics cannot matter and the statistics are bit-reproducible; round 1's fp32 atomics were not, and the
      //          batch-2 image-pooling BatchNorm amplified that into 3-5 % logit differences between runs.  (Beyond that
      //          span the order can move the fp64 sum by one fp64 ulp — invisible after the rounding to fp32 below.) --------
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int e = (warp - 2) * 32 + lane;
        for (int c = e; c < ncols_tile; c += 128) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            a += stat_sm[(w * 2 + 0) * BN + c];
            b += stat_sm[(w * 2 + 1) * BN + c];
          }
          atomicAdd(p.stats + n0 + c, (double)a);
          atomicAdd(p.stats + p.Ncols + n0 + c, (double)b);
        }
        if (p.sync.world > 0) {
          volatile int* flag = reinterpret_cast<volatile int*>(smem_gen + C::STAGES * C::STAGE_BYTES + 128);
          sync_push_when_last(p.sync, p.stats, 2 * p.Ncols, p.stat_ticket, gridDim.x * gridDim.y, e, 128,
                              [] { asm volatile("bar.sync 1, 128;" ::: "memory"); }, flag);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry points resolved at run time; no -lcuda link dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int resolve_driver() {
  static std::once_flag once;
  static int status = 0;
  std::call_once(once, [] {
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q) != cudaSuccess || !f1) status = 1;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q) != cudaSuccess || !f2) status = 1;
    g_encode_tiled = (EncodeTiledFn)f1;
    g_encode_im2col = (EncodeIm2colFn)f2;
    cudaDriverGetVersion(&g_driver_version);
  });
  if (status) set_error("cannot resolve cuTensorMapEncode* driver entry points");
  return status;
}

// 2-D bf16 matrix [rows][cols] with row pitch ld (elements); box = [box_rows][64 cols], 128B swizzle
static int make_map_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  if (resolve_driver()) return 1;
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: base pointer not 16-byte aligned");
  SEG_REQUIRE(ld % 8 == 0, "TMA: row pitch %lld not a multiple of 8 elements", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

// NHWC bf16 tensor, im2col mode: `pixels` base positions x 64 channels per load.  Corners per dimension (w, h).
static int make_map_im2col(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int64_t ld, int lower_w,
                           int lower_h, int upper_w, int upper_h, int stride, int pixels) {
  if (resolve_driver()) return 1;
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: base pointer not 16-byte aligned");
  SEG_REQUIRE(ld % 8 == 0, "TMA: channel pitch %lld not a multiple of 8", (long long)ld);
  SEG_REQUIRE(lower_w >= -128 && upper_w >= -128 && lower_w <= 127 && upper_w <= 127 && lower_h >= -128 &&
                  upper_h >= -128 && lower_h <= 127 && upper_h <= 127,
              "im2col corner out of range");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  int lo[2] = {lower_w, lower_h};
  int up[2] = {upper_w, upper_h};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lo, up,
                               64, (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS,
              "cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d ld=%lld lo=(%d,%d) up=(%d,%d) s=%d", (int)r, N, H,
              W, C, (long long)ld, lower_w, lower_h, upper_w, upper_h, stride);
  // Same workaround CUTLASS applies for drivers <= 13.1: small tensors (< 128 KiB) must clear bit 21 of word 1.
  if (g_driver_version <= 13010) {
    const uint64_t bytes = (uint64_t)N * H * W * ld * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return 0;
}

template <int BN, int STAGES, int KIND>
static int launch_kernel(const TcParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  auto kfn = conv_gemm_tc<BN, STAGES, KIND>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, STAGES>::SMEM);
    SEG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d): %s", Cfg<BN, STAGES>::SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  launch_pdl(kfn, grid, dim3(NTHREADS), (size_t)Cfg<BN, STAGES>::SMEM, stream, p);
  return check_launch("conv_gemm_tc");
}

// Tile configurations: every one fits two CTAs per SM (<= 113 KB smem, <= 256 TMEM columns) so one CTA's epilogue
// overlaps the other's main loop; the staging tile of the epilogue (128 x BN x esize) must fit in STAGES*STAGE_BYTES.
template <int KIND>
static int launch_bn(int bn, const TcParams& p, dim3 grid, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_kernel<64, 4, KIND>(p, grid, stream);    // 4 x 24 KB
    case 128: return launch_kernel<128, 3, KIND>(p, grid, stream);  // 3 x 32 KB
    case 256: return launch_kernel<256, 2, KIND>(p, grid, stream);  // 2 x 48 KB
  }
  set_error("bad BN %d", bn);
  return 1;
}

static int env_dbg() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEG_TC_DBG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace tc
}  // namespace seg
#include "seg_conv_tc2.cuh"
namespace seg {
namespace tc {

static bool long_k_rule() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEG_TC_LONGK");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

static bool use_v2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEG_TC_V2");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

static int env_bn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEG_TC_BN");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static int env_v2bn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEG_TC_V2BN");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// tile width of the persistent kernel for an output of `ncols` columns and `m_tiles` row tiles
// (profiles/conv_sweep_r02.txt: 256-wide tiles win or tie on every C3 shape with >= 256 output columns — 3x3 512->512 d2 @33^2
//  937 -> 1336 TFLOP/s, 1x1 512->2048 722 -> 875, decoder 3x3 256->256 @129^2 1067 -> 1366 — except where the last column
//  block would be mostly padding: dgrad of the 304-channel decoder input, 846 -> 812)
static int pick_v2_bn(int ncols, int64_t m_tiles, bool tma_ok) {
  if (!tma_ok || ncols < 256) return V2_BN;
  const int forced = env_v2bn();
  if (forced == 128 || forced == 256) return forced;
  const int waste = ceil_div(ncols, 256) * 256 - ncols;
  return waste < 128 ? 256 : V2_BN;
}

static int pick_bn(int ncols, int out_dtype) {
  if (ncols <= 64) return 64;
  if (out_dtype == SEG_DT_F32) return 64;  // fp32 staging tile: 128 x 64 x 4 B
  const int forced = env_bn();
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  return 128;
}

bool supported(const seg_conv_desc* d) {
  if (d->C % 8 != 0 || d->ldx % 8 != 0) return false;
  if (d->R != d->S) return false;
  if (d->R * d->S > MAXT) return false;
  if (d->pad > 127 || d->dil * (d->R - 1) - d->pad > 127 || d->dil * (d->R - 1) - d->pad < -127) return false;
  if (d->stride > 8) return false;
  return true;
}

static bool is_pointwise(const seg_conv_desc* d) { return d->R == 1 && d->S == 1 && d->stride == 1 && d->pad == 0; }

int conv_fwd(const seg_conv_desc* d, const void* x, const void* w, void* y, int y_dtype, const float* bias, float beta,
             double* stats, unsigned* stat_ticket, const SyncDesc* sync, cudaStream_t stream) {
  SEG_REQUIRE(supported(d), "tcgen05 conv fwd: unsupported shape (C=%d ldx=%d R=%d pad=%d dil=%d)", d->C, d->ldx, d->R,
              d->pad, d->dil);
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int64_t M = (int64_t)d->N * d->P * d->Q;
  SEG_REQUIRE(M < (1ll << 31), "M too large");
  p.M = (int)M;
  p.Ncols = d->K;
  p.taps = d->R * d->S;
  for (int t = 0; t < p.taps; ++t) {
    p.tap_wt[t] = (short)t;
    p.tap_oh[t] = (short)((t / d->S) * d->dil);
    p.tap_ow[t] = (short)((t % d->S) * d->dil);
  }
  p.kchunks = ceil_div(d->C, BK);
  p.stride = d->stride;
  p.lower_h = p.lower_w = -d->pad;
  p.PQ = d->P * d->Q;
  p.Q = d->Q;
  p.brows_per_tap = d->K;
  p.out = y;
  p.ldo = d->ldy;
  p.out_dtype = y_dtype;
  p.beta = beta;
  p.bias = bias;
  p.stats = stats;
  p.stat_ticket = stat_ticket;
  if (sync && stats) {
    SEG_REQUIRE(stat_ticket != nullptr, "conv fwd: SyncBN needs a zeroed ticket word");
    SEG_REQUIRE(4 * d->K <= sync->n_max, "conv fwd: 2*K = %d fp64 statistics exceed the SyncBN buffer (%d floats)", 2 * d->K, sync->n_max);
    p.sync = *sync;
  }
  const int64_t m_tiles = ceil_div64(M, BM);
  // persistent double-buffered kernel for bf16 outputs wider than 64 channels; the one-tile-per-CTA kernel otherwise
  // (measured: with <= 2 tiles per SM and a long k-loop, two co-resident one-tile CTAs interleave better than one
  //  persistent CTA; everywhere else — short k-loops, many tiles — the persistent kernel wins by 1.3-1.8x)
  const int64_t tiles128 = m_tiles * ceil_div(d->K, 128);
  // (measured, profiles/conv_longk_r02.txt: with few tiles and a long k-loop the two co-resident one-tile CTAs still win for
  //  the 3x3s — 256->256 @33^2: 751 vs 704 TFLOP/s — but not for a 1x1 with >= 256 outputs: 2048->256 612 vs 687 on 256-wide tiles)
  const bool wide_ok = d->K >= 256 && v2_tma_epilogue_ok(y, d->ldy, beta, false);
  const bool long_k_few_tiles = long_k_rule() && tiles128 <= 2 * num_sms() && (int64_t)d->R * d->S * ceil_div(d->C, BK) >= 32 &&
                                !(wide_ok && d->R * d->S == 1);
  const bool v2 = use_v2() && y_dtype == SEG_DT_BF16 && bias == nullptr && d->K > 64 && !long_k_few_tiles;
  // 256-wide persistent tiles (0.75x the operand bytes per flop) where the layer has >= 256 output channels, the output can
  // go through the TMA epilogue and there are enough 128 x 256 tiles to fill the SMs (env SEG_TC_V2BN = 128 | 256 forces)
  const int v2bn = v2 ? pick_v2_bn(d->K, m_tiles, v2_tma_epilogue_ok(y, d->ldy, beta, false)) : V2_BN;
  const int bn = v2 ? v2bn : pick_bn(d->K, y_dtype);
  if (is_pointwise(d)) {
    p.x_im2col = 0;
    if (make_map_2d(&p.mapA, x, M, d->C, d->ldx, BM)) return 1;
  } else {
    p.x_im2col = 1;
    const int upper = d->pad - (d->R - 1) * d->dil;
    if (make_map_im2col(&p.mapA, x, d->N, d->H, d->W, d->C, d->ldx, -d->pad, -d->pad, upper, upper, d->stride, BM)) return 1;
  }
  if (make_map_2d(&p.mapB, w, (int64_t)p.taps * d->K, d->C, d->C, bn)) return 1;
  if (v2) return launch_v2<KIND_KK>(p, stream, v2bn);
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(d->K, bn), 1);
  return launch_bn<KIND_KK>(bn, p, grid, stream);
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int conv_dgrad(const seg_conv_desc* d, const void* dy, const void* w, void* dx, float beta, cudaStream_t stream) {
  SEG_REQUIRE(supported(d) && d->ldy % 8 == 0, "tcgen05 conv dgrad: unsupported shape (stride=%d K=%d ldy=%d)", d->stride,
              d->K, d->ldy);
  const int s = d->stride;
  const int64_t tiles128 = ceil_div64((int64_t)d->N * d->H * d->W, BM) * ceil_div(d->C, 128) / (s * s);
  // (measured: dgrad 256<-256 3x3 @33^2 647 TFLOP/s on the one-tile kernel, 966 on 256-wide persistent tiles)
  const bool wide_ok = s == 1 && pick_v2_bn(d->C, 1 << 20, v2_tma_epilogue_ok(dx, d->ldx, beta, false)) == 256;
  const bool long_k_few_tiles = long_k_rule() && !wide_ok && tiles128 <= 2 * num_sms() &&
                                (int64_t)d->R * d->S * ceil_div(d->K, BK) / (s * s) >= 32;
  const bool v2 = use_v2() && d->C > 64 && !long_k_few_tiles;
  const int bn = v2 ? V2_BN : pick_bn(d->C, SEG_DT_BF16);
  // One launch per parity class (py, px) of the input pixels; stride 1 has the single class (0, 0).
  for (int py = 0; py < s; ++py) {
    for (int px = 0; px < s; ++px) {
      const int Hs = (d->H - py + s - 1) / s, Ws = (d->W - px + s - 1) / s;
      if (Hs <= 0 || Ws <= 0) continue;
      TcParams p;
      memset(&p, 0, sizeof(p));
      // taps whose parity matches: y*1 + pad - r*dil must be a multiple of the stride; offset = that / stride
      int rh[8], oh[8], nh = 0, rw[8], ow[8], nw = 0;
      for (int r = 0; r < d->R; ++r) {
        const int num = py + d->pad - r * d->dil;
        if (((num % s) + s) % s == 0) { rh[nh] = r; oh[nh] = floordiv(num, s); ++nh; }
      }
      for (int q = 0; q < d->S; ++q) {
        const int num = px + d->pad - q * d->dil;
        if (((num % s) + s) % s == 0) { rw[nw] = q; ow[nw] = floordiv(num, s); ++nw; }
      }
      int min_oh = 0, min_ow = 0;
      for (int i = 0; i < nh; ++i) min_oh = (i == 0) ? oh[i] : min(min_oh, oh[i]);
      for (int i = 0; i < nw; ++i) min_ow = (i == 0) ? ow[i] : min(min_ow, ow[i]);
      p.taps = nh * nw;
      for (int i = 0; i < nh; ++i)
        for (int j = 0; j < nw; ++j) {
          const int t = i * nw + j;
          p.tap_wt[t] = (short)(rh[i] * d->S + rw[j]);
          p.tap_oh[t] = (short)(oh[i] - min_oh);
          p.tap_ow[t] = (short)(ow[j] - min_ow);
        }
      const int64_t M = (int64_t)d->N * Hs * Ws;
      SEG_REQUIRE(M < (1ll << 31), "M too large");
      p.M = (int)M;
      p.Ncols = d->C;
      p.kchunks = ceil_div(d->K, BK);
      p.stride = 1;
      p.lower_h = min_oh;
      p.lower_w = min_ow;
      p.PQ = Hs * Ws;
      p.Q = Ws;
      p.brows_per_tap = d->K;
      p.out = dx;
      p.ldo = d->ldx;
      p.out_dtype = SEG_DT_BF16;
      p.beta = beta;
      if (s > 1) {
        p.out_strided = 1;
        p.out_H = d->H;
        p.out_W = d->W;
        p.osy = p.osx = s;
        p.opy = py;
        p.opx = px;
      }
      const int64_t m_tiles = ceil_div64(M, BM);
      if (p.taps > 0) {
        const bool plain = (s == 1 && is_pointwise(d));
        if (plain) {
          p.x_im2col = 0;
          if (make_map_2d(&p.mapA, dy, M, d->K, d->ldy, BM)) return 1;
        } else {
          p.x_im2col = 1;
          // base positions per dimension must number Hs (Ws): box = [lower, dim + upper - 1]
          const int upper_h = Hs - d->P + min_oh, upper_w = Ws - d->Q + min_ow;
          if (make_map_im2col(&p.mapA, dy, d->N, d->P, d->Q, d->K, d->ldy, min_ow, min_oh, upper_w, upper_h, 1, BM)) return 1;
        }
        if (make_map_2d(&p.mapB, w, (int64_t)d->R * d->S * d->K, d->C, d->C, 64)) return 1;
      } else {
        if (beta != 0.f) continue;  // nothing to add to this class
        // no tap reaches this class: the kernel writes zeros (num_iters == 0); maps unused but must be valid objects
        if (make_map_2d(&p.mapA, w, (int64_t)d->R * d->S * d->K, d->C, d->C, BM)) return 1;
        if (make_map_2d(&p.mapB, w, (int64_t)d->R * d->S * d->K, d->C, d->C, 64)) return 1;
      }
      if (v2 && p.taps > 0) {
        const int v2bn = pick_v2_bn(d->C, m_tiles, v2_tma_epilogue_ok(dx, d->ldx, beta, s > 1));
        if (launch_v2<KIND_KM>(p, stream, v2bn)) return 1;
        continue;
      }
      dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(d->C, v2 ? 128 : bn), 1);
      if (launch_bn<KIND_KM>(v2 ? 128 : bn, p, grid, stream)) return 1;
    }
  }
  return 0;
}

int conv_wgrad(const seg_conv_desc* d, const void* dy, const void* x, float* dw, cudaStream_t stream) {
  SEG_REQUIRE(supported(d) && d->ldy % 8 == 0, "tcgen05 conv wgrad: unsupported shape");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int64_t npix = (int64_t)d->N * d->P * d->Q;
  SEG_REQUIRE(npix < (1ll << 31), "too many pixels");
  p.M = d->K;
  p.Ncols = d->C;
  p.taps = d->R * d->S;
  for (int t = 0; t < p.taps; ++t) {
    p.tap_wt[t] = (short)t;
    p.tap_oh[t] = (short)((t / d->S) * d->dil);
    p.tap_ow[t] = (short)((t % d->S) * d->dil);
  }
  p.stride = d->stride;
  p.lower_h = p.lower_w = -d->pad;
  p.PQ = d->P * d->Q;
  p.Q = d->Q;
  p.dw = dw;
  p.dw_K = d->K;
  p.dw_C = d->C;
  p.kblocks_total = (int)ceil_div64(npix, BK);
  static int wgrad_bn_env = -1;
  if (wgrad_bn_env < 0) {
    const char* e = getenv("SEG_TC_WGRAD_BN");
    wgrad_bn_env = e ? atoi(e) : 0;
  }
  // 256-wide wgrad tiles pay on the 3x3 convs with C a multiple of 256 (measured: 2048->256 d12 1231 -> 1376 TFLOP/s, decoder
  // 256->256 @129^2 1182 -> 1332) and lose on the 1x1s (478 -> 363: half as many tiles for the one-wave split-K); env forces
  const bool wide = wgrad_bn_env == 256 ? d->C >= 256 : (wgrad_bn_env == 128 ? false : (d->R * d->S >= 9 && d->C >= 256 && d->C % 256 == 0));
  const int bn = (d->C <= 64) ? 64 : (wide ? 256 : 128);
  const int tiles = ceil_div(d->K, BM) * ceil_div(d->C, bn) * p.taps;
  // one wave: tiles * splits <= resident CTA slots (2 per SM), so no CTA waits for a second wave
  int splits = max(1, (2 * num_sms()) / tiles);
  splits = min(splits, p.kblocks_total);
  splits = min(splits, 1024);
  p.kblocks_per_split = ceil_div(p.kblocks_total, splits);
  splits = ceil_div(p.kblocks_total, p.kblocks_per_split);
  if (make_map_2d(&p.mapA, dy, npix, d->K, d->ldy, 64)) return 1;
  if (is_pointwise(d)) {
    p.x_im2col = 0;
    if (make_map_2d(&p.mapB, x, npix, d->C, d->ldx, 64)) return 1;
  } else {
    p.x_im2col = 1;
    const int upper = d->pad - (d->R - 1) * d->dil;
    if (make_map_im2col(&p.mapB, x, d->N, d->H, d->W, d->C, d->ldx, -d->pad, -d->pad, upper, upper, d->stride, 64)) return 1;
  }
  SEG_REQUIRE((unsigned)(p.taps * splits) <= 65535u, "wgrad grid.z too large");
  dim3 grid((unsigned)ceil_div(d->K, BM), (unsigned)ceil_div(d->C, bn), (unsigned)(p.taps * splits));
  return launch_bn<KIND_MM>(bn, p, grid, stream);
}

}  // namespace tc
}  // namespace seg
