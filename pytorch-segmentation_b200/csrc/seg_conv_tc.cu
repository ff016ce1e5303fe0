// seg_conv_tc.cu — tcgen05 + TMA implicit-GEMM convolution for sm_100a (forward, data-gradient, weight-gradient).
//
// Replaces what the reference dispatches to cuDNN for every nn.Conv2d on the hot path (dense 3x3 / dilated 3x3 /
// 1x1): models/deeplabv3_plus.py:256 (ASPP d=6/12/18), :312-315 (decoder 3x3), torchvision
// Bottleneck conv1/conv2/conv3 (layer4 conv2 dilated at deeplabv3_plus.py:47-53), models/resnet.py:43-48.
//
// One CTA computes a 128 x BN fp32 tile held in TMEM.  Warp roles (192 threads):
//   warp 0  : TMA producer.  The activation operand is fetched with an IM2COL-mode tensor map: one instruction
//             brings 128 output pixels x 64 channels of one filter tap (tap*dilation passed as the im2col offset,
//             padding = the bounding-box corners, out-of-image taps zero-filled by the TMA unit) straight into the
//             128B-swizzled K-major layout tcgen05 consumes — the im2col gather happens in the copy engine, never
//             in HBM.  Weights come through a tiled 2-D map over the packed [tap][K][C] matrix.
//   warp 1  : TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per 64-wide k-block).
//   warps 2-5: epilogue; tcgen05.ld 32 lanes x 32 columns, optional bias / beta-accumulate / per-channel
//             sum & sum-of-squares (BatchNorm statistics fused here), bf16 or fp32 stores.
// Three operand-major combinations of the same pipeline:
//   KK (fprop) : A = activations (K-major),  B = weights  [tap*K + k][c]   (K-major)
//   KM (dgrad) : A = dY im2col  (K-major),  B = weights  [tap*K + k][c]   (MN-major: c contiguous), taps flipped
//   MM (wgrad) : A = dY [pixel][k] (MN-major), B = X im2col [pixel][c] (MN-major), contraction over pixels,
//                split-K over pixel blocks with fp32 atomics into dW[tap][k][c].
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include "seg_common.cuh"
#include "seg_ptx.cuh"

namespace seg {
namespace tc {

using namespace ptx;

constexpr int BM = 128;
constexpr int BK = 64;  // elements per k-block = 128 bytes of bf16
constexpr int A_BYTES = BM * 128;
constexpr int KIND_KK = 0, KIND_KM = 1, KIND_MM = 2;
constexpr int NTHREADS = 192;

struct TcParams {
  CUtensorMap mapA;  // KK/KM: activation-side operand ; MM: dY 2-D
  CUtensorMap mapB;  // KK/KM: packed weights 2-D      ; MM: X (im2col or 2-D)
  int M;             // valid output rows
  int Ncols;         // valid output cols
  int taps, S;
  int kchunks;       // KK/KM: 64-wide channel chunks per tap
  int dil, stride, lower;
  int PQ, Q;         // row index -> (n, p, q)
  int x_im2col;      // activation operand uses the im2col map
  int flip;          // KM: weight tap = taps-1-tap
  int brows_per_tap; // weight-matrix rows per tap
  void* out;
  long long ldo;
  int out_dtype;
  float beta;
  const float* bias;
  float* stats;      // [2*Ncols] or null
  // MM only
  int kblocks_total, kblocks_per_split;
  float* dw;
  int dw_K, dw_C;
};

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void row_to_coords(int m, int PQ, int Q, int stride, int lower, int& n, int& h, int& w) {
  n = m / PQ;
  int rem = m - n * PQ;
  int pp = rem / Q;
  int qq = rem - pp * Q;
  h = lower + pp * stride;
  w = lower + qq * stride;
}

template <int BN, int KIND>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_tc(const __grid_constant__ TcParams p) {
  using C = Cfg<BN>;
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

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0 && num_iters > 0) {
      if (KIND != KIND_MM) {
        int n_img = 0, h0 = 0, w0 = 0;
        if (p.x_im2col) row_to_coords(m0, p.PQ, p.Q, p.stride, p.lower, n_img, h0, w0);
        int it = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          const int wtap = p.flip ? (p.taps - 1 - tap) : tap;
          for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
            const int st = it % C::STAGES;
            const uint32_t ph = (it / C::STAGES) & 1;
            mbar_wait(empty_bar(st), ph ^ 1u);
            const uint32_t a_dst = smem0 + st * C::STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            mbar_arrive_expect_tx(full_bar(st), C::STAGE_BYTES);
            if (p.x_im2col)
              tma_load_im2col_4d(a_dst, &p.mapA, full_bar(st), kc * BK, w0, h0, n_img, (uint16_t)(s * p.dil),
                                 (uint16_t)(r * p.dil));
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
        const int r = tap_mm / p.S, s = tap_mm - r * p.S;
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
            row_to_coords(pix0, p.PQ, p.Q, p.stride, p.lower, n_img, h0, w0);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_im2col_4d(b_dst + j * 8192, &p.mapB, full_bar(st), n0 + j * 64, w0, h0, n_img,
                                 (uint16_t)(s * p.dil), (uint16_t)(r * p.dil));
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
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int row = m0 + lg * 32 + lane;
    if (num_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    float v[32];
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
      if (KIND == KIND_MM) {
        if (row < p.M) {
          float* dst = p.dw + ((size_t)tap_mm * p.dw_K + row) * p.dw_C + col0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.Ncols) atomicAdd(dst + i, v[i]);
        }
        continue;
      }
      const bool row_ok = row < p.M;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.Ncols) v[i] += __ldg(p.bias + col0 + i);
      }
      if (row_ok) {
        const bool full = (col0 + 32 <= p.Ncols);
        if (p.out_dtype == SEG_DT_BF16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
          const bool vec = full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
          if (vec) {
            if (p.beta != 0.f) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                bf16x8 o = reinterpret_cast<const bf16x8*>(dst)[g];
                float f[8];
                unpack8(o, f);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[g * 8 + i] += p.beta * f[i];
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) reinterpret_cast<bf16x8*>(dst)[g] = pack8(v + g * 8);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.Ncols) {
                float o = v[i];
                if (p.beta != 0.f) o += p.beta * bf2f(dst[i]);
                dst[i] = f2bf(o);
              }
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.Ncols) {
              float o = v[i];
              if (p.beta != 0.f) o += p.beta * dst[i];
              dst[i] = o;
            }
        }
      }
      if (p.stats) {
        // column sums over this warp's 32 rows: transpose-reduce, lane i ends with column (col0+i)
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = row_ok ? v[i] : 0.f;
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
        if (col0 + lane < p.Ncols) {
          atomicAdd(p.stats + col0 + lane, s1[0]);
          atomicAdd(p.stats + p.Ncols + col0 + lane, s2[0]);
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

// NHWC bf16 tensor, im2col mode: `pixels` output positions x 64 channels per load
static int make_map_im2col(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int64_t ld, int lower,
                           int upper, int stride, int pixels) {
  if (resolve_driver()) return 1;
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: base pointer not 16-byte aligned");
  SEG_REQUIRE(ld % 8 == 0, "TMA: channel pitch %lld not a multiple of 8", (long long)ld);
  SEG_REQUIRE(lower >= -128 && upper >= -128 && lower <= 127 && upper <= 127, "im2col corner out of range");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  int lo[2] = {lower, lower};
  int up[2] = {upper, upper};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lo, up,
                               64, (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d ld=%lld lo=%d up=%d s=%d",
              (int)r, N, H, W, C, (long long)ld, lower, upper, stride);
  // Same workaround CUTLASS applies for drivers <= 13.1: small tensors (< 128 KiB) must clear bit 21 of word 1.
  if (g_driver_version <= 13010) {
    const uint64_t bytes = (uint64_t)N * H * W * ld * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return 0;
}

template <int BN, int KIND>
static int launch_kernel(const TcParams& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  auto kfn = conv_gemm_tc<BN, KIND>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM);
    SEG_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(smem=%d): %s", Cfg<BN>::SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  kfn<<<grid, NTHREADS, Cfg<BN>::SMEM, stream>>>(p);
  return check_launch("conv_gemm_tc");
}

template <int KIND>
static int launch_bn(int bn, const TcParams& p, dim3 grid, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_kernel<64, KIND>(p, grid, stream);
    case 128: return launch_kernel<128, KIND>(p, grid, stream);
    case 256: return launch_kernel<256, KIND>(p, grid, stream);
  }
  set_error("bad BN %d", bn);
  return 1;
}

static int pick_bn(int ncols, int64_t m_tiles) {
  if (ncols <= 64) return 64;
  if (ncols <= 128) return 128;
  // prefer 256-wide tiles only when they still fill the machine
  const int64_t tiles256 = m_tiles * ceil_div(ncols, 256);
  if (ncols % 256 == 0 && tiles256 >= 2 * num_sms()) return 256;
  return 128;
}

bool supported(const seg_conv_desc* d) {
  if (d->C % 8 != 0 || d->ldx % 8 != 0) return false;
  if (d->R != d->S) return false;
  if (d->pad > 127 || d->dil * (d->R - 1) - d->pad > 127 || d->dil * (d->R - 1) - d->pad < 0) return false;
  if (d->stride > 8) return false;
  return true;
}

static bool is_pointwise(const seg_conv_desc* d) { return d->R == 1 && d->S == 1 && d->stride == 1 && d->pad == 0; }

int conv_fwd(const seg_conv_desc* d, const void* x, const void* w, void* y, int y_dtype, const float* bias, float beta,
             float* stats, cudaStream_t stream) {
  SEG_REQUIRE(supported(d), "tcgen05 conv fwd: unsupported shape (C=%d ldx=%d R=%d pad=%d dil=%d)", d->C, d->ldx, d->R,
              d->pad, d->dil);
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int64_t M = (int64_t)d->N * d->P * d->Q;
  SEG_REQUIRE(M < (1ll << 31), "M too large");
  p.M = (int)M;
  p.Ncols = d->K;
  p.taps = d->R * d->S;
  p.S = d->S;
  p.kchunks = ceil_div(d->C, BK);
  p.dil = d->dil;
  p.stride = d->stride;
  p.lower = -d->pad;
  p.PQ = d->P * d->Q;
  p.Q = d->Q;
  p.flip = 0;
  p.brows_per_tap = d->K;
  p.out = y;
  p.ldo = d->ldy;
  p.out_dtype = y_dtype;
  p.beta = beta;
  p.bias = bias;
  p.stats = stats;
  const int64_t m_tiles = ceil_div64(M, BM);
  const int bn = pick_bn(d->K, m_tiles);
  if (is_pointwise(d)) {
    p.x_im2col = 0;
    if (make_map_2d(&p.mapA, x, M, d->C, d->ldx, BM)) return 1;
  } else {
    p.x_im2col = 1;
    const int upper = d->pad - (d->R - 1) * d->dil;
    if (make_map_im2col(&p.mapA, x, d->N, d->H, d->W, d->C, d->ldx, -d->pad, upper, d->stride, BM)) return 1;
  }
  if (make_map_2d(&p.mapB, w, (int64_t)p.taps * d->K, d->C, d->C, bn)) return 1;
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(d->K, bn), 1);
  return launch_bn<KIND_KK>(bn, p, grid, stream);
}

int conv_dgrad(const seg_conv_desc* d, const void* dy, const void* w, void* dx, float beta, cudaStream_t stream) {
  SEG_REQUIRE(supported(d) && d->stride == 1 && d->ldy % 8 == 0,
              "tcgen05 conv dgrad: unsupported shape (stride=%d K=%d ldy=%d)", d->stride, d->K, d->ldy);
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int64_t M = (int64_t)d->N * d->H * d->W;  // rows = input pixels
  SEG_REQUIRE(M < (1ll << 31), "M too large");
  const int padp = d->dil * (d->R - 1) - d->pad;  // padding of the transposed conv
  // the transposed conv must map P x Q back onto H x W
  SEG_REQUIRE(d->P + 2 * padp - d->dil * (d->R - 1) == d->H && d->Q + 2 * padp - d->dil * (d->S - 1) == d->W,
              "dgrad geometry mismatch");
  p.M = (int)M;
  p.Ncols = d->C;
  p.taps = d->R * d->S;
  p.S = d->S;
  p.kchunks = ceil_div(d->K, BK);
  p.dil = d->dil;
  p.stride = 1;
  p.lower = -padp;
  p.PQ = d->H * d->W;
  p.Q = d->W;
  p.flip = 1;
  p.brows_per_tap = d->K;
  p.out = dx;
  p.ldo = d->ldx;
  p.out_dtype = SEG_DT_BF16;
  p.beta = beta;
  const int64_t m_tiles = ceil_div64(M, BM);
  const int bn = pick_bn(d->C, m_tiles);
  if (is_pointwise(d)) {
    p.x_im2col = 0;
    if (make_map_2d(&p.mapA, dy, M, d->K, d->ldy, BM)) return 1;
  } else {
    p.x_im2col = 1;
    const int upper = padp - (d->R - 1) * d->dil;
    if (make_map_im2col(&p.mapA, dy, d->N, d->P, d->Q, d->K, d->ldy, -padp, upper, 1, BM)) return 1;
  }
  if (make_map_2d(&p.mapB, w, (int64_t)p.taps * d->K, d->C, d->C, 64)) return 1;
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(d->C, bn), 1);
  return launch_bn<KIND_KM>(bn, p, grid, stream);
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
  p.S = d->S;
  p.dil = d->dil;
  p.stride = d->stride;
  p.lower = -d->pad;
  p.PQ = d->P * d->Q;
  p.Q = d->Q;
  p.dw = dw;
  p.dw_K = d->K;
  p.dw_C = d->C;
  p.kblocks_total = (int)ceil_div64(npix, BK);
  const int bn = (d->C <= 64) ? 64 : 128;
  const int tiles = ceil_div(d->K, BM) * ceil_div(d->C, bn) * p.taps;
  int splits = max(1, (2 * num_sms() + tiles - 1) / tiles);
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
    if (make_map_im2col(&p.mapB, x, d->N, d->H, d->W, d->C, d->ldx, -d->pad, upper, d->stride, 64)) return 1;
  }
  SEG_REQUIRE((unsigned)(p.taps * splits) <= 65535u, "wgrad grid.z too large");
  dim3 grid((unsigned)ceil_div(d->K, BM), (unsigned)ceil_div(d->C, bn), (unsigned)(p.taps * splits));
  return launch_bn<KIND_MM>(bn, p, grid, stream);
}

}  // namespace tc
}  // namespace seg
