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
#include "seg_fold.cuh"

namespace seg {


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

// block-level reduction of NACC x 8 per-thread sums over the row lanes; the block's sums become its row of the channel
// slab's fold lane (seg_fold.cuh: fixed-order cross-block sum, bit-reproducible); the block that completes the tree writes
// out[a][c] = beta*out[a][c] + total
template <int NACC>
__device__ __forceinline__ void dw_block_reduce(const DwMap& m, int C, float (*acc)[8], float* fold_rows, unsigned* fold_tickets,
                                                float* __restrict__ out, float beta) {
  __shared__ float red[256 * 8];
  __shared__ int fold_flag;
  const int GB = min(C >> 3, 256);
  const int gl = threadIdx.x % GB;
  const int W = GB * 8;
  const FoldLane L = fold_lane(fold_rows, fold_tickets, blockIdx.y, gridDim.x, NACC * W);
  float* myrow = L.rows1 + (size_t)blockIdx.x * (NACC * W);
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = m.active ? acc[a][i] : 0.f;
    __syncthreads();
    if (m.rl == 0) {
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = 0.f;
      if (m.g < (C >> 3)) {
        for (int r = 0; r < m.rows_par; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] += red[(r * GB + gl) * 8 + i];
      }
      *reinterpret_cast<float4*>(myrow + a * W + gl * 8) = make_float4(s[0], s[1], s[2], s[3]);
      *reinterpret_cast<float4*>(myrow + a * W + gl * 8 + 4) = make_float4(s[4], s[5], s[6], s[7]);
    }
  }
  fold_arrive(L, blockIdx.x, threadIdx.x, 256, [] { __syncthreads(); }, &fold_flag, [&](int c, float v) {
    const int a = c / W, ch = blockIdx.y * W + (c - a * W);
    if (ch < C) {
      float* o = out + (size_t)a * C + ch;
      *o = (beta != 0.f) ? beta * *o + v : v;
    }
  });
}

__global__ void __launch_bounds__(256)
    dwconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ w9, __nv_bfloat16* __restrict__ y,
                      int ldy, int N, int H, int W, int C, int P, int Q, int stride, int pad, int dil,
                      float* __restrict__ stats, float* fold_rows, unsigned* fold_tickets) {
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
      if (stats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += o[i];
          acc[1][i] += o[i] * o[i];
        }
      }
    }
  }
  if (stats) dw_block_reduce<2>(m, C, acc, fold_rows, fold_tickets, stats, 0.f);
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
                             float* fold_rows, unsigned* fold_tickets, float* __restrict__ dw9, float beta) {
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
  dw_block_reduce<9>(m, C, acc, fold_rows, fold_tickets, dw9, beta);
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

static dim3 dw_grid(int64_t M, int C, int blocks_per_sm = 6) {
  const int G = C / 8;
  const int GB = G < 256 ? G : 256;
  const int rows_par = 256 / GB;
  const int gy = ceil_div(G, GB);
  int64_t gx = ceil_div64(M, (int64_t)rows_par * 2);
  const int64_t cap = ((int64_t)num_sms() * blocks_per_sm + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}
// reduction workspace of a dwconv launch: fold rows (floats), then the tickets, in ONE scratch buffer
static int64_t dw_ws_rows(dim3 grid, int C, int nacc) {
  const int G = C / 8;
  const int GB = G < 256 ? G : 256;
  return (int64_t)grid.y * fold_lane_floats((int)grid.x, nacc * GB * 8);
}
static int64_t dw_ws_tickets(dim3 grid) { return (int64_t)grid.y * fold_lane_tickets((int)grid.x); }
constexpr int DW_WGRAD_BLOCKS_PER_SM = 2;  // every block writes a [9][C] row of partial sums: keep the rows few

}  // namespace seg

using namespace seg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

// upper bound (any M) of the scratch a dwconv call needs: fold rows + tickets (seg_fold.cuh)
int64_t seg_dwconv_scratch_floats(int C) {
  const int64_t big = (int64_t)1 << 40;
  const dim3 gf = dw_grid(big, C), gw = dw_grid(big, C, DW_WGRAD_BLOCKS_PER_SM);
  const int64_t f = dw_ws_rows(gf, C, 2) + dw_ws_tickets(gf), w = dw_ws_rows(gw, C, 9) + dw_ws_tickets(gw);
  return (f > w ? f : w) + 64;
}

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
  SEG_REQUIRE(!stats || scratch, "dwconv fwd: stats need a scratch of seg_dwconv_scratch_floats(C) floats");
  const int64_t M = (int64_t)d->N * d->P * d->Q;
  const dim3 grid = dw_grid(M, d->C);
  float* rows = scratch;
  unsigned* tickets = nullptr;
  if (stats) {
    const int64_t nr = (dw_ws_rows(grid, d->C, 2) + 31) / 32 * 32;
    tickets = reinterpret_cast<unsigned*>(scratch + nr);
    cudaMemsetAsync(tickets, 0, (size_t)dw_ws_tickets(grid) * sizeof(unsigned), ST(stream));
  }
  dwconv_fwd_kernel<<<grid, 256, 0, ST(stream)>>>(CBF(x), d->ldx, w9, BF(y), d->ldy, d->N, d->H, d->W, d->C, d->P, d->Q,
                                                  d->stride, d->pad, d->dil, stats, rows, tickets);
  return check_launch("dwconv_fwd");
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
  const dim3 grid = dw_grid(M, d->C, DW_WGRAD_BLOCKS_PER_SM);
  const int64_t nr = (dw_ws_rows(grid, d->C, 9) + 31) / 32 * 32;
  unsigned* tickets = reinterpret_cast<unsigned*>(scratch + nr);
  cudaMemsetAsync(tickets, 0, (size_t)dw_ws_tickets(grid) * sizeof(unsigned), ST(stream));
  dwconv_bwd_weight_kernel<<<grid, 256, 0, ST(stream)>>>(CBF(dy), d->ldy, CBF(x), d->ldx, d->N, d->H, d->W, d->C, d->P, d->Q,
                                                         d->stride, d->pad, d->dil, scratch, tickets, dw9, beta);
  return check_launch("dwconv_bwd_weight");
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
