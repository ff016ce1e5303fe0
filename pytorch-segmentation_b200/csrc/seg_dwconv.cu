// seg_dwconv.cu — depthwise 3x3 (atrous) convolution, the HBM-bound half of the Aligned-Xception backbone
// (SeparableConv2d.conv1, models/deeplabv3_plus.py:77-78: groups = C, stride 1|2, dilation 1|2|4, "same" padding).
// 0.75 % of the network's FLOPs but 348 M activation elements per image: pure bandwidth work, so
// these are vectorised streaming kernels (8 channels = 16 bytes per thread, channel-group-stationary: each thread
// keeps its 9 x 8 filter taps in registers while it strides over pixels) — not GEMMs.
//   fwd       : y = dw(x)            (+ per-channel sum / sum-of-squares of y for the BatchNorm that follows)
//   bwd_data  : dx (+)= dw^T(dy)     (gather form, no atomics)
//   bwd_weight: dw[9][C] += sum_pixels dy * x_shifted   (72 register accumulators per thread, slotted atomics)
// Packed depthwise weights: fp32 [9][C] (tap-major).
#include "seg_common.cuh"

namespace seg {

constexpr int DW_SLOTS = 16;

struct DwMap {
  int g, rl, rows_par;
  bool active;
};
__device__ __forceinline__ DwMap dw_map(int C) {
  const int G = C >> 3;
  const int GB = min(G, 256);
  DwMap r;
  r.rows_par = 256 / GB;
  r.rl = threadIdx.x / GB;
  r.g = blockIdx.y * GB + (threadIdx.x % GB);
  r.active = r.g < G && r.rl < r.rows_par;
  return r;
}

__device__ __forceinline__ void load_taps(const float* __restrict__ w9, int C, int co, float (*w)[8]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    *reinterpret_cast<float4*>(w[t]) = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)t * C + co));
    *reinterpret_cast<float4*>(w[t] + 4) = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)t * C + co + 4));
  }
}

// block-level reduction of NACC x 8 per-thread sums over the row lanes, then slotted atomics into out[slot][NACC][C]
template <int NACC>
__device__ __forceinline__ void dw_block_reduce(const DwMap& m, int C, float (*acc)[8], float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int GB = min(C >> 3, 256);
  const int gl = threadIdx.x % GB;
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = m.active ? acc[a][i] : 0.f;
    __syncthreads();
    if (m.rl == 0 && m.g < (C >> 3)) {
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = 0.f;
      for (int r = 0; r < m.rows_par; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += red[(r * GB + gl) * 8 + i];
      float* o = out + ((size_t)(blockIdx.x % DW_SLOTS) * NACC + a) * C + m.g * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(o + i, s[i]);
    }
  }
}

__global__ void __launch_bounds__(256)
    dwconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ w9, __nv_bfloat16* __restrict__ y,
                      int ldy, int N, int H, int W, int C, int P, int Q, int stride, int pad, int dil,
                      float* __restrict__ stat_slots) {
  const DwMap m = dw_map(C);
  float w[9][8];
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (m.active) {
    const int co = m.g * 8;
    load_taps(w9, C, co, w);
    const int64_t M = (int64_t)N * P * Q;
    const int64_t step = (int64_t)gridDim.x * m.rows_par;
    for (int64_t row = (int64_t)blockIdx.x * m.rows_par + m.rl; row < M; row += step) {
      const int n = (int)(row / (P * Q));
      const int rem = (int)(row - (int64_t)n * P * Q);
      const int p = rem / Q, q = rem - p * Q;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ih = p * stride - pad + r * dil;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int iw = q * stride - pad + s * dil;
          if (iw < 0 || iw >= W) continue;
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(x + (((int64_t)n * H + ih) * W + iw) * ldx + co), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(f[i], w[r * 3 + s][i], o[i]);
        }
      }
      *reinterpret_cast<bf16x8*>(y + row * ldy + co) = pack8(o);
      if (stat_slots) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += o[i];
          acc[1][i] += o[i] * o[i];
        }
      }
    }
  }
  if (stat_slots) dw_block_reduce<2>(m, C, acc, stat_slots);
}

__global__ void __launch_bounds__(256)
    dwconv_bwd_data_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, const float* __restrict__ w9,
                           __nv_bfloat16* __restrict__ dx, int lddx, int N, int H, int W, int C, int P, int Q, int stride,
                           int pad, int dil, float beta) {
  const DwMap m = dw_map(C);
  if (!m.active) return;
  const int co = m.g * 8;
  float w[9][8];
  load_taps(w9, C, co, w);
  const int64_t M = (int64_t)N * H * W;
  const int64_t step = (int64_t)gridDim.x * m.rows_par;
  for (int64_t row = (int64_t)blockIdx.x * m.rows_par + m.rl; row < M; row += step) {
    const int n = (int)(row / (H * W));
    const int rem = (int)(row - (int64_t)n * H * W);
    const int ih = rem / W, iw = rem - ih * W;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int th = ih + pad - r * dil;
      if (th < 0 || (th % stride) != 0) continue;
      const int p = th / stride;
      if (p >= P) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int tw = iw + pad - s * dil;
        if (tw < 0 || (tw % stride) != 0) continue;
        const int q = tw / stride;
        if (q >= Q) continue;
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(dy + (((int64_t)n * P + p) * Q + q) * lddy + co), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(f[i], w[r * 3 + s][i], o[i]);
      }
    }
    __nv_bfloat16* dst = dx + row * lddx + co;
    if (beta != 0.f) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dst), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += beta * f[i];
    }
    *reinterpret_cast<bf16x8*>(dst) = pack8(o);
  }
}

__global__ void __launch_bounds__(256)
    dwconv_bwd_weight_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ x, int ldx,
                             int N, int H, int W, int C, int P, int Q, int stride, int pad, int dil,
                             float* __restrict__ slots /*[DW_SLOTS][9][C]*/) {
  const DwMap m = dw_map(C);
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  if (m.active) {
    const int co = m.g * 8;
    const int64_t M = (int64_t)N * P * Q;
    const int64_t step = (int64_t)gridDim.x * m.rows_par;
    for (int64_t row = (int64_t)blockIdx.x * m.rows_par + m.rl; row < M; row += step) {
      const int n = (int)(row / (P * Q));
      const int rem = (int)(row - (int64_t)n * P * Q);
      const int p = rem / Q, q = rem - p * Q;
      float g[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dy + row * lddy + co), g);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int ih = p * stride - pad + r * dil;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int iw = q * stride - pad + s * dil;
          if (iw < 0 || iw >= W) continue;
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(x + (((int64_t)n * H + ih) * W + iw) * ldx + co), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[r * 3 + s][i] = fmaf(f[i], g[i], acc[r * 3 + s][i]);
        }
      }
    }
  }
  dw_block_reduce<9>(m, C, acc, slots);
}

// sums the slot rows: out[a][c] (=|+=) sum_s slots[s][a][c]   (a < nacc)
__global__ void dw_slot_reduce_kernel(const float* __restrict__ slots, int nacc, int C, float* __restrict__ out, float beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nacc * C) return;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < DW_SLOTS; ++k) s += slots[(size_t)k * nacc * C + i];
  out[i] = (beta != 0.f) ? beta * out[i] + s : s;
}

// depthwise master weight [C][1][3][3] fp32 <-> packed [9][C] fp32
__global__ void dw_pack_kernel(const float* __restrict__ w, float* __restrict__ w9, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  const int t = i / C, c = i - t * C;
  w9[i] = w[(size_t)c * 9 + t];
}
__global__ void dw_unpack_kernel(const float* __restrict__ g9, float* __restrict__ g, int C, float beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  const int c = i / 9, t = i - c * 9;
  const float v = g9[(size_t)t * C + c];
  g[i] = (beta != 0.f) ? beta * g[i] + v : v;
}

static dim3 dw_grid(int64_t M, int C) {
  const int G = C / 8;
  const int GB = G < 256 ? G : 256;
  const int rows_par = 256 / GB;
  const int gy = ceil_div(G, GB);
  int64_t gx = ceil_div64(M, (int64_t)rows_par * 2);
  const int64_t cap = ((int64_t)num_sms() * 6 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

}  // namespace seg

using namespace seg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int64_t seg_dwconv_scratch_floats(int C) { return (int64_t)DW_SLOTS * 9 * C; }

static int dw_check(const seg_conv_desc* d) {
  SEG_REQUIRE(d && d->R == 3 && d->S == 3 && d->K == d->C, "dwconv: 3x3 depthwise (K == C) only");
  SEG_REQUIRE(d->C % 8 == 0 && d->ldx % 8 == 0 && d->ldy % 8 == 0, "dwconv: C / pitches must be multiples of 8");
  const int P = (d->H + 2 * d->pad - d->dil * 2 - 1) / d->stride + 1, Q = (d->W + 2 * d->pad - d->dil * 2 - 1) / d->stride + 1;
  SEG_REQUIRE(P == d->P && Q == d->Q, "dwconv: output size mismatch");
  return 0;
}

int seg_dwconv3x3_fwd(const seg_conv_desc* d, const void* x, const float* w9, void* y, float* stats, float* scratch,
                      void* stream) {
  if (dw_check(d)) return 1;
  SEG_REQUIRE(!stats || scratch, "dwconv fwd: stats need a scratch of 16*2*C floats");
  const int64_t M = (int64_t)d->N * d->P * d->Q;
  if (stats) cudaMemsetAsync(scratch, 0, (size_t)DW_SLOTS * 2 * d->C * sizeof(float), ST(stream));
  dwconv_fwd_kernel<<<dw_grid(M, d->C), 256, 0, ST(stream)>>>(CBF(x), d->ldx, w9, BF(y), d->ldy, d->N, d->H, d->W, d->C, d->P, d->Q,
                                                              d->stride, d->pad, d->dil, stats ? scratch : nullptr);
  if (check_launch("dwconv_fwd")) return 1;
  if (stats) {
    dw_slot_reduce_kernel<<<ceil_div(2 * d->C, 128), 128, 0, ST(stream)>>>(scratch, 2, d->C, stats, 1.0f);
    return check_launch("dw_slot_reduce");
  }
  return 0;
}

int seg_dwconv3x3_bwd_data(const seg_conv_desc* d, const void* dy, const float* w9, void* dx, float beta, void* stream) {
  if (dw_check(d)) return 1;
  const int64_t M = (int64_t)d->N * d->H * d->W;
  dwconv_bwd_data_kernel<<<dw_grid(M, d->C), 256, 0, ST(stream)>>>(CBF(dy), d->ldy, w9, BF(dx), d->ldx, d->N, d->H, d->W, d->C,
                                                                   d->P, d->Q, d->stride, d->pad, d->dil, beta);
  return check_launch("dwconv_bwd_data");
}

int seg_dwconv3x3_bwd_weight(const seg_conv_desc* d, const void* dy, const void* x, float* dw9, float beta, float* scratch,
                             void* stream) {
  if (dw_check(d)) return 1;
  SEG_REQUIRE(scratch != nullptr, "dwconv bwd_weight: scratch of seg_dwconv_scratch_floats(C) floats required");
  const int64_t M = (int64_t)d->N * d->P * d->Q;
  cudaMemsetAsync(scratch, 0, (size_t)DW_SLOTS * 9 * d->C * sizeof(float), ST(stream));
  dwconv_bwd_weight_kernel<<<dw_grid(M, d->C), 256, 0, ST(stream)>>>(CBF(dy), d->ldy, CBF(x), d->ldx, d->N, d->H, d->W, d->C,
                                                                     d->P, d->Q, d->stride, d->pad, d->dil, scratch);
  if (check_launch("dwconv_bwd_weight")) return 1;
  dw_slot_reduce_kernel<<<ceil_div(9 * d->C, 128), 128, 0, ST(stream)>>>(scratch, 9, d->C, dw9, beta);
  return check_launch("dw_slot_reduce");
}

int seg_dw_pack_weight(const float* w_c133, float* w9, int C, void* stream) {
  dw_pack_kernel<<<ceil_div(9 * C, 128), 128, 0, ST(stream)>>>(w_c133, w9, C);
  return check_launch("dw_pack");
}
int seg_dw_unpack_wgrad(const float* g9, float* g_c133, int C, float beta, void* stream) {
  dw_unpack_kernel<<<ceil_div(9 * C, 128), 128, 0, ST(stream)>>>(g9, g_c133, C, beta);
  return check_launch("dw_unpack");
}

}  // extern "C"
