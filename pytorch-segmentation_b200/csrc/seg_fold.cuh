// seg_fold.cuh — deterministic cross-block reduction ("ticket tree") used wherever a per-channel sum crosses CTAs:
// BatchNorm batch statistics from the conv epilogues (utils/sync_batchnorm/batchnorm.py:128-145 / nn.BatchNorm2d's
// sum and sum of squares), the BN backward sums, the depthwise-conv statistics.
//
// Round 1 accumulated these with fp32 atomics, whose order varies from run to run; at batch 2 the image-pooling
// BatchNorm amplified that last-bit noise into 3-5 % logit differences between two runs of the same step.  Now every
// contributing block writes its partial sums to ITS OWN row with plain stores and takes a ticket; the last block of a
// group of FOLD_G rows adds the group's rows in row order (in fp64) into a level-2 row, and the last group to finish
// adds the level-2 rows in group order.  Which block happens to be last does not matter: the additions and their order
// are fixed, so the totals are bit-identical from run to run (and independent of the SM schedule).
//
// A "lane" is one independent reduction domain (a conv column block, a streaming kernel's channel slab): `nrows`
// contributing blocks x `width` floats.  Workspace per lane: rows1[nrows][width] + rows2[ngroups][width] floats (need
// not be initialised) and ngroups + 1 uint32 tickets that MUST be zero at launch.
#pragma once
#include <stdint.h>

namespace seg {

constexpr int FOLD_G = 32;

__host__ __device__ inline int fold_groups(int nrows) { return (nrows + FOLD_G - 1) / FOLD_G; }
// floats of row workspace per lane / tickets per lane
__host__ __device__ inline int64_t fold_lane_floats(int nrows, int width) { return (int64_t)(nrows + fold_groups(nrows)) * width; }
__host__ __device__ inline int fold_lane_tickets(int nrows) { return fold_groups(nrows) + 1; }

struct FoldLane {
  float* rows1;       // [nrows][width]
  float* rows2;       // [ngroups][width]
  unsigned* tickets;  // [ngroups + 1], zero at launch
  int nrows, width;
};

__device__ __forceinline__ FoldLane fold_lane(float* rows_ws, unsigned* tickets_ws, int lane, int nrows, int width) {
  FoldLane L;
  L.rows1 = rows_ws + (size_t)lane * fold_lane_floats(nrows, width);
  L.rows2 = L.rows1 + (size_t)nrows * width;
  L.tickets = tickets_ws + (size_t)lane * fold_lane_tickets(nrows);
  L.nrows = nrows;
  L.width = width;
  return L;
}

// Sum of `n` (<= FOLD_G... any) floats spaced `stride` apart, added in index order in fp64.  The loads of a chunk of 16 are
// all issued before the first add (a dependent-load chain here cost 20-50 us of kernel tail in the first version).
__device__ __forceinline__ double fold_column(const float* base, int n, size_t stride) {
  double s = 0.0;
  for (int r0 = 0; r0 < n; r0 += 16) {
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = (r0 + r < n) ? __ldcg(base + (size_t)(r0 + r) * stride) : 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += (double)v[r];
  }
  return s;
}

// Called by the `nthr` cooperating threads of a block (tid 0..nthr-1; `sync()` is a barrier over exactly those threads)
// AFTER they have written row `row` of L.rows1 with plain stores.  In exactly one block of the lane — the one that
// completes the tree — it returns true after calling emit(column, total) for every column; everywhere else false.
// `sm_flag`: a shared-memory word private to the cooperating threads.
template <class Sync, class Emit>
__device__ __forceinline__ bool fold_arrive(const FoldLane& L, int row, int tid, int nthr, Sync sync, volatile int* sm_flag,
                                            Emit emit) {
  const int ngroups = fold_groups(L.nrows);
  const int grp = row / FOLD_G;
  const int g0 = grp * FOLD_G;
  const int gn = min(FOLD_G, L.nrows - g0);
  __threadfence();  // this thread's row stores are visible device-wide before the ticket is taken
  sync();
  if (tid == 0) *sm_flag = (atomicAdd(L.tickets + grp, 1u) == (unsigned)(gn - 1));
  sync();
  if (!*sm_flag) return false;
  __threadfence();
  if (ngroups == 1) {
    for (int c = tid; c < L.width; c += nthr) emit(c, (float)fold_column(L.rows1 + (size_t)g0 * L.width + c, gn, L.width));
    return true;
  }
  for (int c = tid; c < L.width; c += nthr)
    L.rows2[(size_t)grp * L.width + c] = (float)fold_column(L.rows1 + (size_t)g0 * L.width + c, gn, L.width);
  __threadfence();
  sync();
  if (tid == 0) *sm_flag = (atomicAdd(L.tickets + ngroups, 1u) == (unsigned)(ngroups - 1));
  sync();
  if (!*sm_flag) return false;
  __threadfence();
  for (int c = tid; c < L.width; c += nthr) emit(c, (float)fold_column(L.rows2 + c, ngroups, L.width));
  return true;
}

}  // namespace seg
