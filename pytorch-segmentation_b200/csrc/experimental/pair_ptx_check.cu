// Compile-only translation unit: instantiates every wrapper of seg_ptx_pair.cuh so that ptxas (sm_100a) validates the
// instruction forms.  Never launched, never linked into libseg_b200.so (tools/check_pair_ptx.sh).
#include "seg_ptx_pair.cuh"

using namespace seg::ptx2;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_ptx_check(const __grid_constant__ CUtensorMap map2d, const __grid_constant__ CUtensorMap map4d, uint32_t* out) {
  __shared__ __align__(1024) uint8_t tile[16384];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * i), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    tmem2_alloc((uint32_t)__cvta_generic_to_shared(&tmem_slot), 256);
    tmem2_relinquish();
  }
  __syncthreads();
  cluster_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 32) {
    tma2_load_2d(tile_a, &map2d, bar0, 0, (int)rank * 128);
    tma2_load_im2col_4d(tile_a + 8192, &map4d, bar0, 0, 0, 0, 0, (uint16_t)1, (uint16_t)1);
    tma_load_2d_multicast(tile_a, &map2d, bar0 + 8, 0, 0, (uint16_t)0x3);
  }
  if (threadIdx.x == 64 && rank == 0) {
    umma2_bf16(tmem, 0ull, 0ull, make_idesc_bf16_m256(256, 0, 0), 0u);
    umma2_commit_multicast(bar0 + 16, (uint16_t)0x3);
    umma_commit_multicast(bar0 + 24, (uint16_t)0x3);
  }
  if (threadIdx.x == 96) mbar_arrive_cta0(bar0 + 16);
  cluster_sync();
  if (threadIdx.x < 32) tmem2_dealloc(tmem, 256);
  if (threadIdx.x == 0) out[blockIdx.x] = tmem;
}
